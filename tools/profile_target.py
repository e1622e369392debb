#!/usr/bin/env python
"""Minimal launch sequence for ncu: a few fused fwd+bwd launches at the benchmark size in both native layouts."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "lagrange-dual-learning_b200")):
  if p not in sys.path:
    sys.path.insert(0, p)
import torch
from ikl import ConstraintSpec, standard_config, fused, synthetic

B = 1 << int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 24
layouts = sys.argv[2].split(",") if len(sys.argv) > 2 else ["planes", "rows"]
cfg_name = sys.argv[3] if len(sys.argv) > 3 else "cfg_joint_limits_and_4obs_dynamic"
mode = sys.argv[4] if len(sys.argv) > 4 else "fwdbwd"        # "grad": the gradient-only launch of a training step
spec = ConstraintSpec.from_config(standard_config(cfg_name), 3, 0.005, 1.0)
theta, q_des, obst = synthetic.make_batch(B, seed=0, device="cuda:0")
lam = list(synthetic.SHIPPED_LAMBDA_OBS4_DYN)[:spec.num_constraints]
for layout in layouts:
  th, tg, ob = synthetic.to_planes(theta, q_des, obst) if layout == "planes" else (theta, q_des, obst)
  dth = torch.empty_strided(th.shape, th.stride(), dtype=torch.float32, device="cuda:0")
  for _ in range(3):
    if mode == "grad":
      fused.lagrangian_backward(spec, th, tg, ob if spec.dynamic else None, lam, dtheta=dth)
      l = dth[0, 0]
    else:
      v, l, _ = fused.lagrangian_forward_backward(spec, th, tg, ob if spec.dynamic else None, lam, dtheta=dth)
  torch.cuda.synchronize()
  print(layout, float(l))
