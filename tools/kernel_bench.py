#!/usr/bin/env python
"""Kernel-only timing table for the fused Lagrangian kernels (CUDA events, back-to-back launches, inputs >> L2).

  python tools/kernel_bench.py [--batch-log2 24] [--iters 30] [--libs a.so,b.so]   # libs: variant builds to compare

Each lib is measured in a fresh subprocess (IKL_LIB) so variants do not share state.  Prints one JSON line per row.
"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "lagrange-dual-learning_b200")
for p in (ROOT, PKG):
  if p not in sys.path:
    sys.path.insert(0, p)


def measure(args):
  import torch
  from ikl import ConstraintSpec, standard_config, fused, synthetic, _cabi
  dev = torch.device("cuda:0")
  B = 1 << args.batch_log2
  J = args.num_links
  lengths = tuple(float(v) for v in args.link_lengths.split(",")) if args.link_lengths else None
  theta, q_des, obst = synthetic.make_batch(B, seed=0, device="cuda:0", num_links=J)
  planes = synthetic.to_planes(theta, q_des, obst)
  rows = (theta, q_des, obst)
  peak = 6548.8
  try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
  except Exception:
    pass
  # algorithmic bytes per sample (generalised to J links): theta 4J + target 8 (+ 4 obstacles x 12) in, dtheta 4J out
  dyn_f, stat_f = 4 * J + 8 + 48, 4 * J + 8
  cases = [("cfg_joint_limits_and_4obs_dynamic", dyn_f + 4 * J, dyn_f), ("cfg_joint_limits_and_4obs_static", stat_f + 4 * J, stat_f),
           ("cfg_joint_limits_and_1obs_static", stat_f + 4 * J, stat_f), ("cfg_joint_limits", stat_f + 4 * J, stat_f),
           ("cfg_no_constraints", stat_f + 4 * J, stat_f)]
  if args.quick:
    cases = cases[:1]
  if args.configs:
    cases = [c for c in cases if any(c[0].endswith(n) for n in args.configs.split(","))]
  for cfg_name, bytes_fb, bytes_f in cases:
    spec = ConstraintSpec.from_config(standard_config(cfg_name), J, 0.005, 1.0, link_lengths=lengths)
    K = spec.num_constraints
    lam = [1.0 + 0.1 * i for i in range(K)]
    lam_dev = torch.tensor(lam, device=dev)
    for lname, (th, tg, ob) in (("planes", planes), ("rows", rows)):
      if args.layouts and lname not in args.layouts.split(","):
        continue
      dth = torch.empty_strided(th.shape, th.stride(), dtype=torch.float32, device=dev)
      for mode in ("fwdbwd_hostw", "fwdbwd_devw", "fwd", "grad_hostw", "grad_devw"):
        if args.quick and mode in ("fwdbwd_devw", "grad_devw"):
          continue
        def call():
          if mode == "fwd":
            return fused.lagrangian_forward(spec, th, tg, ob if spec.dynamic else None, lam)
          if mode.startswith("grad_"):      # gradient-only launch (ikl_lagrangian_backward): d/dtheta, no sums
            return fused.lagrangian_backward(spec, th, tg, ob if spec.dynamic else None, lam if mode == "grad_hostw" else lam_dev, dtheta=dth)
          return fused.lagrangian_forward_backward(spec, th, tg, ob if spec.dynamic else None, lam if mode == "fwdbwd_hostw" else lam_dev, dtheta=dth)
        for _ in range(3):
          call()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e9
        tot = 0.0
        for rep in range(3):
          e0.record()
          for _ in range(args.iters):
            call()
          e1.record()
          torch.cuda.synchronize()
          ms = e0.elapsed_time(e1) / args.iters
          best = min(best, ms)
          tot += ms
        nbytes = bytes_f if mode == "fwd" else bytes_fb
        print(json.dumps({"lib": os.path.basename(_cabi.loaded_path()), "config": cfg_name, "links": J, "layout": lname, "mode": mode,
                          "algorithmic_bytes_per_sample": nbytes,
                          "kernel": _cabi.last_kernel_name(), "ms_mean": round(tot / 3, 4), "ms_best": round(best, 4),
                          "gsamples_per_s": round(B / best / 1e6, 2), "algo_GBps": round(nbytes * B / best / 1e6, 1),
                          "frac_of_hbm_peak": round(nbytes * B / best / 1e6 / peak, 4)}), flush=True)


def measure_elementwise(args):
  """The function-level drop-in kernels (level (i)): pure streaming maps, timed the same way."""
  import torch
  from ikl import ops, synthetic, _cabi
  B = 1 << args.batch_log2
  theta, q_des, obst = synthetic.make_batch(B, seed=0, device="cuda:0")
  peak = 6548.8
  try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
  except Exception:
    pass
  lim = torch.tensor([-2.094395, 2.094395], device="cuda:0").unsqueeze(0).expand(B, -1)
  # (name, algorithmic bytes per sample, bytes per sample in the 32-byte sectors the strided views touch, call)
  cases = [
      ("fk_forward (12 B in, 36 B out per sample)", 48, 48, lambda: ops.fk_forward_kinematics(theta)),
      ("circle_penalty on column views of (B,3) and (B,4,4) tensors (5 x 4 B in, 4 B out)", 24, 12 + 32 + 4,
       lambda: ops.circle_penalty(q_des[:, 0], q_des[:, 1], obst[:, 0, 0], obst[:, 0, 1], obst[:, 0, 2], 1.0, 0.005)),
      ("joint_limit_penalty on a column view of (B,3) (4 B in, 4 B out)", 8, 12 + 4,
       lambda: ops.joint_limit_penalty(theta[:, 1], lim[:, 0], lim[:, 1], 0.005, 1.0)),
  ]
  for name, nbytes, sector_bytes, fn in cases:
    for _ in range(3):
      fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.iters):
      fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.iters
    print(json.dumps({"lib": os.path.basename(_cabi.loaded_path()), "kernel": name, "ms_mean": round(ms, 4),
                      "gsamples_per_s": round(B / ms / 1e6, 2), "algo_GBps": round(nbytes * B / ms / 1e6, 1),
                      "frac_of_hbm_peak": round(nbytes * B / ms / 1e6 / peak, 4),
                      "sector_GBps": round(sector_bytes * B / ms / 1e6, 1),
                      "frac_of_hbm_peak_on_sector_bytes": round(sector_bytes * B / ms / 1e6 / peak, 4),
                      "note": "includes the torch allocation of the output; strided column views move whole 32-byte sectors"}), flush=True)


def measure_callers(args):
  """Rows f-1 / f-2 of SURVEY 8: on-device dataset generation and evaluation statistics (the steps either side of the path)."""
  import torch
  from ikl import ConstraintSpec, standard_config, synthetic, device_data, _cabi
  peak = 6548.8
  try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
  except Exception:
    pass
  B = 1 << args.batch_log2
  ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  for cfg_name in ("cfg_joint_limits_and_4obs_dynamic", "cfg_joint_limits_and_4obs_static"):
    cfg = standard_config(cfg_name)
    d = device_data.generate_dataset(cfg, B, seed=1, device="cuda:0")
    torch.cuda.synchronize()
    ev0.record()
    for rep in range(3):
      del d                                   # hand the output blocks back to the caching allocator: no cudaMalloc in the timed region
      d = device_data.generate_dataset(cfg, B, seed=2 + rep, device="cuda:0")
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / 3
    out_bytes = 24 + (64 if d["random_obstacles"] is not None else 0)       # theta 12 + ee 12 (+ obstacle rows 64) written per example
    print(json.dumps({"lib": os.path.basename(_cabi.loaded_path()), "kernel": "dataset_generate (f-1) " + cfg_name, "ms_mean": round(ms, 4),
                      "gsamples_per_s": round(B / ms / 1e6, 3), "algo_GBps": round(out_bytes * B / ms / 1e6, 1),
                      "frac_of_hbm_peak": round(out_bytes * B / ms / 1e6 / peak, 4), "exhausted": int(d["exhausted"].item()),
                      "note": "Philox + FK + rejection sampling per example, includes the torch allocation of the outputs; "
                              "reference: 6.2 s per 8000 examples on the host (SURVEY 8f-1)"}), flush=True)
    del d
  theta, q_des, obst = synthetic.make_batch(B, seed=0, device="cuda:0")
  spec = ConstraintSpec.from_config(standard_config("cfg_joint_limits_and_4obs_dynamic"), 3, 0.005, 1.0)
  for lname, (th, tg, ob) in (("planes", synthetic.to_planes(theta, q_des, obst)), ("rows", (theta, q_des, obst))):
    counts, pos = device_data.eval_stats(spec, th, tg, ob)
    torch.cuda.synchronize()
    ev0.record()
    for _ in range(args.iters):
      device_data.eval_stats(spec, th, tg, ob, counts=counts, pos_err_sum=pos)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / args.iters
    print(json.dumps({"lib": os.path.basename(_cabi.loaded_path()), "kernel": "eval_stats (f-2) " + lname, "ms_mean": round(ms, 4),
                      "gsamples_per_s": round(B / ms / 1e6, 2), "algo_GBps": round(68 * B / ms / 1e6, 1),
                      "frac_of_hbm_peak": round(68 * B / ms / 1e6 / peak, 4),
                      "note": "per-constraint violation counts + any-JL / any-obstacle rates + position error; 68 algorithmic bytes per sample"}),
          flush=True)


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--batch-log2", type=int, default=24)
  ap.add_argument("--iters", type=int, default=30)
  ap.add_argument("--libs", default="")
  ap.add_argument("--quick", action="store_true")
  ap.add_argument("--configs", default="", help="comma-separated config-name suffixes to keep, e.g. 4obs_static,no_constraints")
  ap.add_argument("--layouts", default="", help="planes and/or rows")
  ap.add_argument("--num-links", type=int, default=3, choices=[2, 3])
  ap.add_argument("--link-lengths", default="", help="comma-separated link lengths (default: unit lengths)")
  ap.add_argument("--repeat", type=int, default=1, help="with --libs: cycle through the list this many times (take the best per lib)")
  ap.add_argument("--child", action="store_true")
  args = ap.parse_args()
  if args.child or not args.libs:
    measure(args)
    if not args.quick and not args.configs and not args.layouts and args.num_links == 3 and not args.link_lengths:
      measure_elementwise(args)
      measure_callers(args)
    return
  for lib in args.libs.split(",") * max(1, args.repeat):     # --repeat: A B C A B C ... so that clock drift hits every lib alike
    env = dict(os.environ, IKL_LIB=os.path.join(PKG, "lib", lib) if not os.path.isabs(lib) else lib)
    cmd = [sys.executable, os.path.abspath(__file__), "--child", "--batch-log2", str(args.batch_log2), "--iters", str(args.iters)]
    if args.quick:
      cmd.append("--quick")
    if args.configs:
      cmd += ["--configs", args.configs]
    if args.layouts:
      cmd += ["--layouts", args.layouts]
    cmd += ["--num-links", str(args.num_links)]
    if args.link_lengths:
      cmd += ["--link-lengths", args.link_lengths]
    subprocess.run(cmd, env=env, check=False)


if __name__ == "__main__":
  main()
