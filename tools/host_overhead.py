#!/usr/bin/env python
"""Host-side cost of one fused launch through the C ABI at a tiny batch (where it is all that matters): issue rate of
back-to-back asynchronous calls, per layout, with and without the TMA-staged kernel (tensor maps are encoded per call).

  python tools/host_overhead.py [--batch 8] [--calls 5000]
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "lagrange-dual-learning_b200")):
  if p not in sys.path:
    sys.path.insert(0, p)


def child(args):
  import torch
  from ikl import ConstraintSpec, standard_config, fused, synthetic, _cabi
  spec = ConstraintSpec.from_config(standard_config("cfg_joint_limits_and_4obs_dynamic"), 3, 0.005, 1.0)
  theta, q_des, obst = synthetic.make_batch(args.batch, seed=0, device="cuda:0")
  lam = [1.0] * 16
  for layout in ("rows", "planes"):
    th, tg, ob = synthetic.to_planes(theta, q_des, obst) if layout == "planes" else (theta, q_des, obst)
    dth = torch.empty_strided(th.shape, th.stride(), dtype=torch.float32, device="cuda:0")
    for _ in range(200):
      fused.lagrangian_forward_backward(spec, th, tg, ob, lam, dtheta=dth)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.calls):
      fused.lagrangian_forward_backward(spec, th, tg, ob, lam, dtheta=dth)
    t_issue = time.perf_counter() - t0
    torch.cuda.synchronize()
    t_all = time.perf_counter() - t0
    print(json.dumps({"layout": layout, "kernel": _cabi.last_kernel_name(), "batch": args.batch, "tma_env": os.environ.get("IKL_PLANES_TMA", ""),
                      "us_per_call_issue": round(1e6 * t_issue / args.calls, 2), "us_per_call_complete": round(1e6 * t_all / args.calls, 2)}), flush=True)


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--batch", type=int, default=8)
  ap.add_argument("--calls", type=int, default=5000)
  ap.add_argument("--child", action="store_true")
  args = ap.parse_args()
  if args.child:
    return child(args)
  for tma in ("1", "0"):
    subprocess.run([sys.executable, os.path.abspath(__file__), "--child", "--batch", str(args.batch), "--calls", str(args.calls)],
                   env=dict(os.environ, IKL_PLANES_TMA=tma), check=False)


if __name__ == "__main__":
  main()
