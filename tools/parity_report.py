#!/usr/bin/env python
"""Measured parity of the CUDA path against the float64 oracle at full benchmark size -- numbers, not pass/fail.

    python tools/parity_report.py [--batch-log2 24] [--out profiles/r02/parity_report.json]

For each of the five task configs, on a synthetic batch of 2^batch_log2 samples (ikl.synthetic.make_batch), in the
planes and the rows layout:
  positions   function-level FK drop-in vs float64: max and 99.99th-percentile |delta|, absolute and relative to the arm's reach (3)
  penalties   function-level circle / joint-limit penalties (per sample) vs float64 on a 2^22 slice: max |delta| relative to max |ref|
  dtheta      fused kernel vs the float64 closed-form gradient: max and 99.99th percentile of the per-sample max-norm error
              relative to S = max |ref|, (a) over all samples, (b) away from kinks (oracle |term| >= 1e-6), plus the fraction
              of samples masked as near a kink and the fraction of unmasked samples whose error exceeds 1e-5 * S (the ones
              that need the conditioning term 3e-6 * sum_i |w_i| slope / d_i of tests/test_gpu_parity.py::grad_close)
  sums        the K constraint sums vs float64, relative to sum_b |term|
  lambda      12 dual-ascent steps (ikl_dual_update, lr 1e-4) on 12 different batches vs the same recursion in float64
The float64 truth comes from oracle/ik_oracle.c (OpenMP; agrees with oracle/ik_oracle.py to 1e-12, tests/test_oracle_pinned.py).
Checker tooling: it may import oracle/ (it is not part of the product tree).
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "lagrange-dual-learning_b200")
for p in (ROOT, PKG):
  if p not in sys.path:
    sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

CONFIGS = ("cfg_no_constraints", "cfg_joint_limits", "cfg_joint_limits_and_1obs_static", "cfg_joint_limits_and_4obs_static",
           "cfg_joint_limits_and_4obs_dynamic")
GOOD, BAD = 0.005, 1.0


def pct(a, q):
  """q-th percentile of a large 1-D array without a full sort."""
  k = min(len(a) - 1, int(np.ceil(q / 100.0 * len(a))) - 1)
  return float(np.partition(a, k)[k])


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--batch-log2", type=int, default=24)
  ap.add_argument("--out", default="")
  args = ap.parse_args()
  from ikl import ConstraintSpec, standard_config, fused, synthetic, ops, _cabi
  from oracle import build_oracle as C, ik_oracle as O
  dev = torch.device("cuda:0")
  B = 1 << args.batch_log2
  report = {"batch": B, "slopes": [GOOD, BAD], "oracle": "oracle/ik_oracle.c float64 (OpenMP, %d threads)" % C.max_threads(),
            "gpu": torch.cuda.get_device_name(0), "configs": {}}
  theta, q_des, obst = synthetic.make_batch(B, seed=2024, device="cuda:0")
  th_h, qd_h, ob_h = theta.cpu().numpy(), q_des.cpu().numpy(), obst.cpu().numpy()

  # ---- positions: the FK drop-in (same for every config) -------------------------------------------------------------
  t0 = time.time()
  diag = C.diagnostics_f64(O.spec_from_json(standard_config(CONFIGS[-1]), 3, GOOD, BAD), th_h, qd_h, ob_h, np.ones(16))
  tips = ops.fk_forward_kinematics(theta)
  got = torch.stack([tips[0][:, 0], tips[0][:, 1], tips[1][:, 0], tips[1][:, 1], tips[2][:, 0], tips[2][:, 1]], dim=1).cpu().numpy()
  err = np.abs(got.astype(np.float64) - diag["tips"]).max(axis=1)
  report["positions"] = {"max_abs": float(err.max()), "p99_99_abs": pct(err, 99.99), "max_rel_to_reach_3": float(err.max() / 3.0),
                         "kernel": "ikl_fk_forward (function-level drop-in)"}
  del tips, got, err

  # ---- per-sample penalties: function-level drop-ins on a 2^22 slice -------------------------------------------------
  n = min(B, 1 << 22)
  os5 = O.spec_from_json(standard_config(CONFIGS[-1]), 3, GOOD, BAD)
  terms = C.diagnostics_f64(os5, th_h[:n], qd_h[:n], ob_h[:n], np.ones(16), want_terms=True)["terms"]
  tips_n = ops.fk_forward_kinematics(theta[:n])
  lim = torch.tensor([-2.094395, 2.094395], device=dev).unsqueeze(0).expand(n, -1)
  worst_c = worst_j = 0.0
  for j in range(3):
    pj = ops.joint_limit_penalty(theta[:n, j], lim[:, 0], lim[:, 1], GOOD, BAD).cpu().numpy().astype(np.float64)
    worst_j = max(worst_j, float(np.abs(pj - terms[:, 1 + j]).max() / np.abs(terms[:, 1 + j]).max()))
  for o in range(4):
    for j in range(3):
      pc = ops.circle_penalty(tips_n[j][:, 0], tips_n[j][:, 1], obst[:n, o, 0], obst[:n, o, 1], obst[:n, o, 2], BAD, GOOD).cpu().numpy()
      ref = terms[:, 4 + 3 * o + j]
      worst_c = max(worst_c, float(np.abs(pc.astype(np.float64) - ref).max() / np.abs(ref).max()))
  report["penalties"] = {"samples": n, "joint_limit_max_rel_to_max_ref": worst_j, "circle_max_rel_to_max_ref": worst_c,
                         "note": "circle penalties evaluated on the fp32 tips of the FK drop-in, like scripts/trainer.py:233-238"}
  del terms, tips_n

  # ---- fused kernel: dtheta and sums, every config x layout ----------------------------------------------------------
  for name in CONFIGS:
    cfg = standard_config(name)
    spec, ospec = ConstraintSpec.from_config(cfg, 3, GOOD, BAD), O.spec_from_json(cfg, 3, GOOD, BAD)
    K = spec.num_constraints
    lam = (1.0 + 0.25 * np.arange(K)).astype(np.float32)
    sums64, abs_sums, dth64 = C.lagrangian_f64(ospec, th_h, qd_h, ob_h, lam.astype(np.float64))
    d = C.diagnostics_f64(ospec, th_h, qd_h, ob_h, lam.astype(np.float64))
    kink = d["kink_dist"] < 1e-6
    S = float(np.abs(dth64).max())
    entry = {"K": K, "grad_scale_S": S, "kink_fraction": float(kink.mean())}
    for layout in ("planes", "rows"):
      th, tg, ob = synthetic.to_planes(theta, q_des, obst, max(spec.num_obstacles, 1)) if layout == "planes" else (theta, q_des, obst)
      viol, loss, dth = fused.lagrangian_forward_backward(spec, th, tg, ob if spec.dynamic else None, lam.tolist())
      kname = _cabi.last_kernel_name()
      e = np.abs(dth.cpu().numpy().astype(np.float64) - dth64).max(axis=1)
      keep = ~kink
      needs_cond = keep & (e > 1e-5 * S)
      bound = 1e-5 * S + 3e-6 * d["cond"]
      over = keep & (e > bound)
      rel_sums = np.abs(viol.cpu().numpy().astype(np.float64) - sums64) / np.maximum(abs_sums, 1e-300)
      entry[layout] = {
          "kernel": kname,
          "dtheta_max_rel_all": float(e.max() / S), "dtheta_p99_99_rel_all": pct(e, 99.99) / S,
          "dtheta_max_rel_away_from_kinks": float(e[keep].max() / S), "dtheta_p99_99_rel_away_from_kinks": pct(e[keep], 99.99) / S,
          "fraction_above_1e-5_S_away_from_kinks": float(needs_cond.sum() / max(1, keep.sum())),
          "fraction_above_conditioning_bound": float(over.sum() / max(1, keep.sum())),
          "sums_max_rel_to_sum_abs_terms": float(rel_sums.max()),
          "loss_rel": float(abs(float(loss) - float((lam * sums64).sum())) / float((lam * abs_sums).sum()))}
      del dth, e
    report["configs"][name] = entry
    print(name, json.dumps(entry), flush=True)

  # ---- lambda after 12 dual steps -------------------------------------------------------------------------------------
  cfg = standard_config(CONFIGS[-1])
  spec, ospec = ConstraintSpec.from_config(cfg, 3, GOOD, BAD), O.spec_from_json(cfg, 3, GOOD, BAD)
  lam_g = torch.ones(16, device=dev)
  lam_o = np.ones(16)
  nb = 1 << 20
  for step in range(12):
    th, qd, ob = synthetic.make_batch(nb, seed=9000 + step, device="cuda:0")
    accum = torch.zeros(16, dtype=torch.float64, device=dev)
    fused.lagrangian_forward(spec, th, qd, ob, lam_g.tolist(), viol_accum=accum)
    lam_g = fused.dual_update(lam_g, accum, 1e-6)
    s64, _, _ = C.lagrangian_f64(ospec, th.cpu().numpy(), qd.cpu().numpy(), ob.cpu().numpy(), lam_o, want_grad=False)
    lam_o = np.maximum(0.0, lam_o + 1e-6 * s64)
  report["lambda_after_12_dual_steps"] = {"max_rel_err": float(np.abs(lam_g.cpu().numpy() - lam_o).max() / np.abs(lam_o).max()),
                                          "batch_per_step": nb, "multiplier_lr": 1e-6, "lambda_max": float(lam_o.max())}
  report["seconds"] = time.time() - t0
  text = json.dumps(report, indent=1)
  print(text)
  if args.out:
    os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
    with open(args.out, "w") as f:
      f.write(text + "\n")


if __name__ == "__main__":
  main()
