"""Evaluation statistics of a trained IK network (reference: scripts/evaluation/visualize_ik_solutions.py:50-119, minus
the meshcat drawing): average end-effector position error, per-constraint violation counts (> 1e-4), and how often any
joint limit / any obstacle constraint is violated -- the numbers behind Table 1 of the reference's report.

The reference pushes the validation set through process_batch one example at a time and thresholds the per-example
constraint vector on the host; here the whole set goes through the network in large batches and ikl_eval_stats does
the thresholding and counting on the device.

  python lagrange-dual-learning_b200/fused_overlay/scripts/evaluation/evaluate_ik_solutions.py --model_name NAME --config_file CFG --load_weights_folder DIR [...]
"""
import os
import sys

import torch

_OVERLAY = os.path.abspath(os.path.join(os.path.dirname(os.path.realpath(__file__)), "..", ".."))
for _p in (os.path.dirname(_OVERLAY), _OVERLAY):        # package root (ikl, kinematics, utils, models), then this overlay (scripts)
  if _p not in sys.path:
    sys.path.insert(0, _p)

from ikl.device_data import eval_stats  # noqa: E402
from scripts.options import Options  # noqa: E402
from scripts.trainer import IkLagrangeDualTrainer  # noqa: E402


def evaluate(trainer, batch_size=65536, use_groundtruth_theta=False):
  """Returns a dict: position_err (mean of sqrt(sum(position_err_sq))), num_violations (K,), any_joint_limit,
  any_obstacle (counts) and n (number of validation examples)."""
  ds = trainer.val_dataset
  n = len(ds)
  dev = trainer.device
  was_training = trainer.model.training
  trainer.model.eval()                      # the reference leaves dropout on (no model.eval() anywhere); evaluation should not
  counts = torch.zeros(18, dtype=torch.int64, device=dev)
  err = torch.zeros(1, dtype=torch.float64, device=dev)
  template = ds.tensor_template.to(dev)
  with torch.no_grad():
    for s in range(0, n, batch_size):
      e = min(n, s + batch_size)
      ee = ds.random_ee[s:e].to(dev)
      x = template.unsqueeze(0).repeat(e - s, 1)
      x[:, 0:2] = ee[:, :2]
      obst = None
      if ds.random_obstacles is not None:
        obst = ds.random_obstacles[s:e].to(dev)
        x[:, 4:] = obst.reshape(e - s, 16)
      theta = ds.random_theta[s:e].to(dev) if use_groundtruth_theta else trainer.model(x)
      eval_stats(trainer.spec, theta.contiguous(), ee, obst if trainer.spec.dynamic else None, 1e-4, counts, err)
  if was_training:
    trainer.model.train()
  K = trainer.spec.num_constraints
  counts = counts.cpu()
  return {"n": n, "position_err": float(err.item()) / max(n, 1), "num_violations": counts[:K].tolist(),
          "any_joint_limit": int(counts[16]), "any_obstacle": int(counts[17])}


def report(trainer, stats):
  n = max(stats["n"], 1)
  for i, name in enumerate(trainer.constraint_names):
    print("{} | {} | {}%".format(name, stats["num_violations"][i], float(stats["num_violations"][i]) / n))
  print("Avg position error =", stats["position_err"])
  print("Joint limit (%) =", float(stats["any_joint_limit"]) / n)
  print("Obstacle (%)=", float(stats["any_obstacle"]) / n)


if __name__ == "__main__":
  opt = Options().parse()
  trainer = IkLagrangeDualTrainer(opt)
  report(trainer, evaluate(trainer, use_groundtruth_theta=opt.show_groundtruth_theta))
