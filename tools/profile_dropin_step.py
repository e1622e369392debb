#!/usr/bin/env python
"""Where the time of level (i) goes: the reference's UNMODIFIED IkLagrangeDualTrainer.process_batch + backward()
(baseline/_ref/scripts/trainer.py, loaded by file path) over this repo's function-level CUDA drop-ins, profiled with
torch.profiler.  python tools/profile_dropin_step.py [output file]"""
import contextlib
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "lagrange-dual-learning_b200")):
  if p not in sys.path:
    sys.path.insert(0, p)
import torch
from torch.profiler import profile, ProfilerActivity
from ikl import synthetic, _cabi
from baseline import ref_harness as H

CONFIG_NAME = "cfg_joint_limits_and_4obs_dynamic"
dev = torch.device("cuda:0")
trainer_mod, options_mod = H.load_reference_scripts()
H.assert_resolution(trainer_mod)
tmp = tempfile.mkdtemp(prefix="ikl_prof_dropin_")
opt = H.reference_options(options_mod, ["--model_name", "prof", "--config_file", os.path.join(H.REF_TREE, "resources", CONFIG_NAME + ".json"),
                                        "--log_dir", tmp, "--num_workers", "0", "--train_dataset_size", "16", "--val_dataset_size", "16",
                                        "--batch_size", "16", "--penalty_good_slope", "0.005", "--penalty_bad_slope", "1.0", "--hidden_units", "16"])
with H.git_worktree(tmp), contextlib.redirect_stdout(sys.stderr):
  tr = trainer_mod.IkLagrangeDualTrainer(opt)
lam_dev = torch.tensor(list(synthetic.SHIPPED_LAMBDA_OBS4_DYN), device=dev)


class FixedTheta(torch.nn.Module):
  def __init__(self, t):
    super().__init__()
    self.theta = torch.nn.Parameter(t.clone())

  def forward(self, x):
    return self.theta


lines = []
for lg in (16, 20, 22):
  B = 1 << lg
  theta, q_des, obst = synthetic.make_batch(B, seed=0, device="cuda:0")
  tr.model = FixedTheta(theta).to(dev)
  inputs = {"input_tensor": torch.zeros(B, 20, device=dev), "q_ee_desired": q_des, "joint_theta": theta.clone(), "dynamic_obstacles": obst}

  def step():
    out = tr.process_batch(dict(inputs), lam_dev)
    tr.model.zero_grad()
    out["lagrange_loss"].backward()
  step()
  torch.cuda.synchronize()
  n0 = _cabi.launch_count()
  t0 = time.perf_counter()
  for _ in range(3):
    step()
  t_issue = (time.perf_counter() - t0) / 3
  torch.cuda.synchronize()
  t_done = (time.perf_counter() - t0) / 3
  lines.append("batch 2^%d: %.2f ms per step (host returns after %.2f ms), %d libikl launches per step" % (
      lg, t_done * 1e3, t_issue * 1e3, (_cabi.launch_count() - n0) // 3))
  if lg == 22:
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
      for _ in range(3):
        step()
      torch.cuda.synchronize()
    lines.append(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=60))
    lines.append(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=12, max_name_column_width=60))
text = "\n".join(lines)
print(text)
if len(sys.argv) > 1:
  with open(sys.argv[1], "w") as f:
    f.write("level (i): the reference's unmodified process_batch + backward() over the CUDA drop-ins, %s, config %s\n" % (
        torch.cuda.get_device_name(0), CONFIG_NAME))
    f.write(text + "\n")
