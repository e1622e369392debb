#!/usr/bin/env python
"""Launch each function-level drop-in kernel a few times at the benchmark size (for `ncu --metrics gpu__time_duration.sum`
or `--set full -k regex:...`): FK forward/backward, circle penalty forward/backward, joint-limit penalty forward/backward on
the trainer's strided column views (scripts/trainer.py:201-204, :233-238 of the reference)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "lagrange-dual-learning_b200")):
  if p not in sys.path:
    sys.path.insert(0, p)
import torch
from ikl import synthetic
from kinematics.forward_kinematics import ForwardKinematicsThreeLinkTorch
from kinematics.penalties import piecewise_circle_penalty, piecewise_joint_limit_penalty

B = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 24)
theta, q_des, obst = synthetic.make_batch(B, seed=0, device="cuda:0")
theta.requires_grad_(True)
lim = torch.tensor([[-2.094395, 2.094395]], device="cuda:0").expand(B, -1)
for _ in range(3):
  q_all = ForwardKinematicsThreeLinkTorch(theta)
  g_jl = piecewise_joint_limit_penalty(theta[:, 1], lim[:, 0], lim[:, 1], 0.005, 1.0).sum()
  g_ob = piecewise_circle_penalty(q_all[0][:, 0], q_all[0][:, 1], obst[:, 0, 0], obst[:, 0, 1], obst[:, 0, 2], 1.0, 0.005).sum()
  (g_jl + g_ob).backward()
  theta.grad = None
torch.cuda.synchronize()
print("ok", float(g_jl), float(g_ob))
