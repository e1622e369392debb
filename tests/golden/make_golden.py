#!/usr/bin/env python
"""Generate the golden fixtures in tests/golden/ by RUNNING THE UNMODIFIED REFERENCE.

Run from the repo root, in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

The reference (miloknowles/lagrange-dual-learning) is imported straight from /root/reference:
kinematics/forward_kinematics.py, kinematics/penalties.py and the real
scripts/trainer.py:IkLagrangeDualTrainer.  Two shims make trainer.py importable here without
touching it: a stub ``tensorboardX`` module (absent in this image) and cwd inside a git work
tree for ``git.Repo(search_parent_directories=True)`` (trainer.py:105).

Nothing here is read at test time except the .npz/.pt files this script writes.  The GPU box has
no /root/reference, which is exactly why these vectors are committed.
"""
import argparse
import json
import math
import os
import sys
import tempfile
import types

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))

SHIPPED_LAMBDA = {  # resources/pretrained_models/*/models/weights_*/multipliers.pth (fp32, CPU-mapped)
    "cfg_no_constraints": "ik_threelink_position/models/weights_600",
    "cfg_joint_limits": "ik_threelink_joint_limits/models/weights_200",
    "cfg_joint_limits_and_1obs_static": "ik_threelink_obs1/models/weights_950",
    "cfg_joint_limits_and_4obs_static": "ik_threelink_obs4/models/weights_950",
    "cfg_joint_limits_and_4obs_dynamic": "ik_threelink_obs4_dyn/models/weights_650",
}
CONFIGS = list(SHIPPED_LAMBDA.keys())


def import_reference():
  sys.path.insert(0, REF)
  stub = types.ModuleType("tensorboardX")

  class SummaryWriter(object):          # accepts trainer.py:116's logdir= and swallows everything
    def __init__(self, *a, **k): pass
    def add_scalar(self, *a, **k): pass
    def add_scalars(self, *a, **k): pass
    def close(self): pass
  stub.SummaryWriter = SummaryWriter
  sys.modules["tensorboardX"] = stub
  import kinematics.forward_kinematics as fk
  import kinematics.penalties as pen
  import scripts.trainer as trainer_mod
  import scripts.options as options_mod
  import models.simple_network as net_mod
  return fk, pen, trainer_mod, options_mod, net_mod


def gen_inputs(B, seed, dynamic=True):
  """Same distribution as the benchmark's synthetic batches, tiny B."""
  g = torch.Generator().manual_seed(seed)
  theta = (torch.rand(B, 3, generator=g) * 2 - 1) * 2.5
  theta_t = (torch.rand(B, 3, generator=g) * 2 - 1) * 2.094395
  obst = torch.zeros(B, 4, 4)
  obst[:, :, :2] = (torch.rand(B, 4, 2, generator=g) * 2 - 1) * 1.2
  obst[:, :, 2] = 0.4
  return theta, theta_t, obst


def golden_fk(fk):
  g = torch.Generator().manual_seed(11)
  theta = (torch.rand(249, 3, generator=g) * 2 - 1) * 4.0
  special = torch.tensor([
      [0.0, 0.0, 0.0],
      [math.pi / 2, -math.pi / 2, -math.pi / 2],          # scripts/misc/visualize_example.py:16
      [math.pi, math.pi, math.pi],
      [-2.094395, 2.094395, -2.094395],
      [100.0, -37.5, 12.25],                                # large-argument range reduction
      [1e-8, -1e-8, 1e-20],
      [1000.0, 2000.0, -3000.0],
      [3.1415927, -3.1415927, 1.5707964],
  ])
  theta = torch.cat([special, theta])
  out32 = fk.ForwardKinematicsThreeLinkTorch(theta)
  out64 = fk.ForwardKinematicsThreeLinkTorch(theta.double())
  # 2-link known answers from the NumPy scalar FK (forward_kinematics.py:7-22)
  th2 = theta[:32, :2].double().numpy()
  ee2 = np.stack([np.asarray(fk.ForwardKinematicsTwoLink(t), dtype=np.float64) for t in th2])
  np.savez(os.path.join(HERE, "fk.npz"), theta=theta.numpy(),
           x3=out32[0].numpy(), x2=out32[1].numpy(), x1=out32[2].numpy(),
           x3_f64=out64[0].numpy(), x2_f64=out64[1].numpy(), x1_f64=out64[2].numpy(),
           theta2=th2, ee2_f64=ee2)


def golden_penalties(pen):
  g = torch.Generator().manual_seed(12)
  N = 384
  jx = (torch.rand(N, generator=g) * 2 - 1) * 3
  jy = (torch.rand(N, generator=g) * 2 - 1) * 3
  ox = (torch.rand(N, generator=g) * 2 - 1) * 1.2
  oy = (torch.rand(N, generator=g) * 2 - 1) * 1.2
  rad = torch.rand(N, generator=g) * 0.6 + 0.1
  # edge rows: exactly on the circle (+x, +y, 3-4-5 triangle), far away, just inside, on centre is
  # kept separate (its reference gradient is NaN)
  e_jx = torch.tensor([0.75, 0.5, 0.8, -1.0, 0.5 + 0.2499, 0.0])
  e_jy = torch.tensor([0.5, 0.75, 0.9, 0.5, 0.5, 0.0])
  e_ox = torch.tensor([0.5, 0.5, 0.5, 0.5, 0.5, 0.5])
  e_oy = torch.tensor([0.5, 0.5, 0.5, 0.5, 0.5, 0.5])
  e_r = torch.tensor([0.25, 0.25, 0.5, 0.25, 0.25, 0.25])
  jx, jy = torch.cat([e_jx, jx]), torch.cat([e_jy, jy])
  ox, oy, rad = torch.cat([e_ox, ox]), torch.cat([e_oy, oy]), torch.cat([e_r, rad])
  slopes = [(1.0, 0.1), (1.0, 0.005), (0.1, 1.0), (0.5, 0.5), (0.0, 0.3)]
  out = {"jx": jx.numpy(), "jy": jy.numpy(), "ox": ox.numpy(), "oy": oy.numpy(), "radius": rad.numpy(),
         "circle_slopes": np.array(slopes)}
  for i, (a_in, a_out) in enumerate(slopes):
    for dt, tag in ((torch.float32, ""), (torch.float64, "_f64")):
      x = jx.to(dt).clone().requires_grad_(True)
      y = jy.to(dt).clone().requires_grad_(True)
      p = pen.piecewise_circle_penalty(x, y, ox.to(dt), oy.to(dt), rad.to(dt), inside_slope=a_in, outside_slope=a_out)
      p.sum().backward()
      out["circle_%d%s" % (i, tag)] = p.detach().numpy()
      out["circle_%d_djx%s" % (i, tag)] = x.grad.numpy()
      out["circle_%d_djy%s" % (i, tag)] = y.grad.numpy()
  # the centre-of-circle case: value is finite, reference gradient is NaN
  x = torch.tensor([0.5], requires_grad=True)
  y = torch.tensor([0.5], requires_grad=True)
  p = pen.piecewise_circle_penalty(x, y, torch.tensor([0.5]), torch.tensor([0.5]), torch.tensor([0.25]), 1.0, 0.1)
  p.sum().backward()
  out["circle_centre_value"] = p.detach().numpy()
  out["circle_centre_djx"] = x.grad.numpy()

  # joint limits: symmetric (+-2pi/3 as float32 of the JSON literal), asymmetric, degenerate
  th = (torch.rand(N, generator=g) * 2 - 1) * 3.5
  lims = [(-2.094395, 2.094395), (-0.5, 1.75), (-3.14159, 3.14159), (1.0, -1.0)]
  jl_slopes = [(0.1, 1.0), (0.005, 1.0), (1.0, 0.1), (0.25, 0.25)]
  out["jl_limits"] = np.array(lims)
  out["jl_slopes"] = np.array(jl_slopes)
  for li, (lo, hi) in enumerate(lims):
    lo_t, hi_t = torch.Tensor([lo]), torch.Tensor([hi])
    mid = ((lo_t + hi_t) / 2.0).item()
    half = (0.5 * torch.abs(hi_t - lo_t)).item()
    edge = torch.tensor([mid, lo_t.item(), hi_t.item(), mid + half, mid - half, mid + 1e-3, lo - 1.0, hi + 1.0], dtype=torch.float32)
    tt = torch.cat([edge, th])
    out["jl_theta_%d" % li] = tt.numpy()
    for si, (a_in, a_out) in enumerate(jl_slopes):
      for dt, tag in ((torch.float32, ""), (torch.float64, "_f64")):
        t = tt.to(dt).clone().requires_grad_(True)
        p = pen.piecewise_joint_limit_penalty(t, lo_t.to(dt), hi_t.to(dt), inside_slope=a_in, outside_slope=a_out)
        p.sum().backward()
        out["jl_%d_%d%s" % (li, si, tag)] = p.detach().numpy()
        out["jl_%d_%d_dtheta%s" % (li, si, tag)] = t.grad.numpy()
  np.savez(os.path.join(HERE, "penalties.npz"), **out)


class FixedTheta(torch.nn.Module):
  """Stands in for trainer.model: returns a preset theta leaf so process_batch's FK/penalty part is isolated."""
  def __init__(self, theta):
    super().__init__()
    self.theta = torch.nn.Parameter(theta.clone())

  def forward(self, x):
    return self.theta


def make_opt(options_mod, cfg_path, log_dir, **over):
  argv = ["--model_name", "golden", "--config_file", cfg_path, "--log_dir", log_dir, "--num_workers", "0",
          "--train_dataset_size", "16", "--val_dataset_size", "16", "--batch_size", "16",
          "--penalty_good_slope", "0.005", "--penalty_bad_slope", "1.0", "--hidden_units", "16",
          "--initial_lambda", "1", "--optimizer_lr", "5e-5", "--multiplier_lr", "1e-4"]
  for k, v in over.items():
    argv += ["--" + k, str(v)]
  o = options_mod.Options()
  return o.parser.parse_args(argv)


def golden_process_batch(trainer_mod, options_mod, tmp):
  B = 96
  for ci, name in enumerate(CONFIGS):
    cfg_path = os.path.join(REF, "resources", name + ".json")
    opt = make_opt(options_mod, cfg_path, tmp)
    tr = trainer_mod.IkLagrangeDualTrainer(opt)
    theta, theta_t, obst = gen_inputs(B, seed=100 + ci)
    # a few hand-placed rows: exactly at a joint limit / mid-range, EE exactly on target
    theta[0] = torch.tensor([0.3, -0.5, 0.8])
    theta[1] = torch.tensor([2.5, -1.0, 0.2])
    theta[2] = torch.tensor([0.0, 0.0, 0.0])
    import kinematics.forward_kinematics as fk
    q_des = fk.ForwardKinematicsThreeLinkTorch(theta_t)[0].clone()
    q_des[2] = torch.tensor([3.0, 0.0, 0.0])
    K = len(tr.lambda_multipliers)
    lam_shipped = torch.load(os.path.join(REF, "resources/pretrained_models", SHIPPED_LAMBDA[name], "multipliers.pth"),
                             map_location="cpu")
    assert lam_shipped.shape[0] == K, (name, lam_shipped.shape, K)
    lam_ramp = 1.0 + 0.25 * torch.arange(K, dtype=torch.float32)
    out = {"theta": theta.numpy(), "q_ee_desired": q_des.numpy(), "dynamic_obstacles": obst.numpy(),
           "lam_shipped": lam_shipped.numpy(), "lam_ramp": lam_ramp.numpy(),
           "good_slope": np.float64(opt.penalty_good_slope), "bad_slope": np.float64(opt.penalty_bad_slope),
           "constraint_names": np.array(tr.constraint_names)}
    for tag, lam in (("shipped", lam_shipped), ("ramp", lam_ramp)):
      tr.model = FixedTheta(theta)
      inputs = {"input_tensor": torch.zeros(B, 20), "q_ee_desired": q_des.clone(), "joint_theta": theta_t.clone()}
      if tr.json_config["dynamic_constraints"]["random_obstacles"]:
        inputs["dynamic_obstacles"] = obst.clone()
      o = tr.process_batch(inputs, lam.clone())
      tr.model.zero_grad()
      o["lagrange_loss"].backward()
      out["viol_" + tag] = o["constraint_violations"].detach().numpy()
      out["loss_" + tag] = o["lagrange_loss"].detach().numpy()
      out["dtheta_" + tag] = tr.model.theta.grad.detach().numpy().copy()
      out["q_ee_actual"] = o["q_ee_actual"].detach().numpy()
      out["position_err_sq"] = o["position_err_sq"].detach().numpy()
      out["joint_angle_l2"] = o["joint_angle_l2"].detach().numpy()
    np.savez(os.path.join(HERE, "process_batch_%s.npz" % name), **out)


def golden_dual_trajectory(trainer_mod, options_mod, net_mod, tmp):
  """Real trainer.main() for a few dual epochs with dropout removed (p=0) so it is device-independent.

  One batch == the whole (tiny) dataset, so DataLoader shuffling only permutes rows inside the sum.
  Saves: the datasets in the reference's own cache format (kinematics/dataset.py:45-50), the initial
  state_dict, lambda after every epoch, and an isolated update_multipliers() call over 4 val batches.
  """
  for name in ("cfg_joint_limits_and_4obs_dynamic", "cfg_joint_limits_and_1obs_static"):
    cfg_path = os.path.join(REF, "resources", name + ".json")
    n_epochs = 6
    opt = make_opt(options_mod, cfg_path, tmp, train_dataset_size=64, val_dataset_size=32, batch_size=64,
                   lagrange_iters=n_epochs, train_iters=1, optimizer_lr=1e-3, multiplier_lr=1e-2, model_save_hz=1000)
    torch.manual_seed(1234)
    tr = trainer_mod.IkLagrangeDualTrainer(opt)
    tr.model = net_mod.SimpleNetwork(20, 3, hidden_units=opt.hidden_units, depth=8, dropout=0.0)
    tr.optimizer = torch.optim.Adam(tr.model.parameters(), lr=opt.optimizer_lr, betas=(0.9, 0.999))
    init_state = {k: v.clone() for k, v in tr.model.state_dict().items()}
    lam_hist = [tr.lambda_multipliers.clone()]
    for epoch in range(n_epochs):          # trainer.py:327-333 without validate()'s printing
      tr.train_with_relaxation(tr.lambda_multipliers)
      tr.lambda_multipliers = tr.update_multipliers(tr.lambda_multipliers)
      lam_hist.append(tr.lambda_multipliers.clone())
    final_state = {k: v.clone() for k, v in tr.model.state_dict().items()}

    # isolated dual step over 4 validation batches of 8 with the final model
    from torch.utils.data import DataLoader
    tr.val_loader = DataLoader(tr.val_dataset, 8, False, num_workers=0)
    lam_iso = tr.update_multipliers(lam_hist[0].clone())

    def ds_dict(ds):
      return {"random_theta": ds.random_theta, "random_ee": ds.random_ee,
              "random_obstacles": ds.random_obstacles, "tensor_template": ds.tensor_template}
    torch.save({"train": ds_dict(tr.train_dataset), "val": ds_dict(tr.val_dataset),
                "init_state": init_state, "final_state": final_state,
                "lambda_history": torch.stack(lam_hist), "lambda_isolated_step": lam_iso,
                # log_dir is a random temporary directory and the trainer records the work tree's commit (trainer.py:105-107):
                # both are stored as placeholders so that the file is byte-reproducible
                "opt": {k: ("<tmp>" if k == "log_dir" else "<head>" if k.startswith("commit_has") else v) for k, v in vars(opt).items()
                        if isinstance(v, (int, float, str, bool, type(None)))}},
               os.path.join(HERE, "dual_trajectory_%s.pt" % name))


def main():
  ap = argparse.ArgumentParser()
  ap.parse_args()
  assert os.path.isdir(REF), "needs the read-only reference mount"
  torch.set_num_threads(1)               # fixed reduction tree => reproducible fixture bits
  fk, pen, trainer_mod, options_mod, net_mod = import_reference()
  golden_fk(fk)
  golden_penalties(pen)
  with tempfile.TemporaryDirectory() as tmp:
    golden_process_batch(trainer_mod, options_mod, tmp)
    golden_dual_trajectory(trainer_mod, options_mod, net_mod, tmp)
  meta = {"torch": torch.__version__, "numpy": np.__version__, "reference": REF,
          "note": "generated by tests/golden/make_golden.py from the unmodified reference"}
  with open(os.path.join(HERE, "META.json"), "w") as f:
    json.dump(meta, f, indent=1)
  print("golden fixtures written to", HERE)


if __name__ == "__main__":
  main()
