#!/usr/bin/env python
"""Kernel-level breakdown of one large-batch train step (torch.profiler): SimpleNetwork forward (cuBLAS), the fused
Lagrangian, backward, Adam -- the micro-step bench.py's train_step metric repeats (planes data, last layer transposed,
no gradient rescaling pass).  python tools/profile_train_step.py [log2 batch=20] [output file]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "lagrange-dual-learning_b200")):
  if p not in sys.path:
    sys.path.insert(0, p)
import torch
from torch.profiler import profile, ProfilerActivity
from ikl import ConstraintSpec, standard_config, fused, synthetic
from models.simple_network import SimpleNetwork

B = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 20)
dev = torch.device("cuda:0")
spec = ConstraintSpec.from_config(standard_config("cfg_joint_limits_and_4obs_dynamic"), 3, 0.005, 1.0)
lam = list(synthetic.SHIPPED_LAMBDA_OBS4_DYN)
torch.manual_seed(1234)
model = SimpleNetwork(20, 3, hidden_units=100, depth=8, dropout=0.05).to(dev)
opt = torch.optim.Adam(model.parameters(), lr=5e-5, betas=(0.9, 0.999))
theta0, q_des, obst = synthetic.make_batch(B, seed=1000, device="cuda:0")
x = torch.zeros(B, 20, device=dev)
x[:, 0:2] = q_des[:, :2]
x[:, 2], x[:, 3] = -synthetic.LIMIT, synthetic.LIMIT
x[:, 4:] = obst.reshape(B, 16)
_, q_pl, ob_pl = synthetic.to_planes(theta0, q_des, obst)
lam_dev = torch.tensor(lam, device=dev)


def step():
  theta = model.forward_planes(x)
  viol, loss = fused.FusedIkLagrangian.apply(theta, q_pl, ob_pl, lam_dev, spec, lam, True)
  model.zero_grad()
  loss.backward()
  opt.step()


for _ in range(3):
  step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
  for _ in range(3):
    step()
  torch.cuda.synchronize()
table = prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70)
print(table)
if len(sys.argv) > 2:
  with open(sys.argv[2], "w") as f:
    f.write("torch.profiler, 3 train steps of %d samples (SimpleNetwork 20->100x8->3 fp32 on cuBLAS, fused IK Lagrangian, Adam) on %s\n" % (
        B, torch.cuda.get_device_name(0)))
    f.write(table + "\n")
