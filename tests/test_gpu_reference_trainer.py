"""The north_star acceptance sentence, executed: the reference's OWN scripts/trainer.py (byte-identical copy in
baseline/_ref, loaded by file path) runs on the B200 with `kinematics.*`, `utils.*`, `models.*` resolving to this repo's CUDA
drop-ins -- op by op through the function-level kernels, with the reference's own CopySlices pattern
(`viol[j] = penalty(...).sum()`, trainer.py:198-204, :228-240), its CPU-built static obstacle table (:218-226) and its
DataLoader.  Results are compared with the golden outputs the same trainer produced on the CPU with the reference's own
kinematics (tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

from conftest import CONFIG_NAMES, GOLDEN, golden_npz
from baseline import ref_harness as H
from test_gpu_parity import RTOL, close, grad_close, oracle_truth, specs_for

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not H.have_reference_tree(), reason="baseline/_ref not installed (python tools/install_ref.py)")]


class FixedTheta(torch.nn.Module):
  """Stands in for trainer.model (as in make_golden.py): returns a preset theta leaf, isolating the FK / penalty part."""

  def __init__(self, theta):
    super().__init__()
    self.theta = torch.nn.Parameter(theta.clone())

  def forward(self, x):
    return self.theta


def make_reference_trainer(name, tmp_path, **over):
  trainer_mod, options_mod = H.load_reference_scripts()
  H.assert_resolution(trainer_mod)
  assert trainer_mod.device == "cuda"
  cfg_path = os.path.join(H.REF_TREE, "resources", name + ".json")           # the reference's own config files, unchanged
  argv = ["--model_name", "golden", "--config_file", cfg_path, "--log_dir", str(tmp_path), "--num_workers", "0",
          "--train_dataset_size", "16", "--val_dataset_size", "16", "--batch_size", "16", "--penalty_good_slope", "0.005",
          "--penalty_bad_slope", "1.0", "--hidden_units", "16", "--initial_lambda", "1", "--optimizer_lr", "5e-5",
          "--multiplier_lr", "1e-4"]
  for k, v in over.items():
    argv += ["--" + k, str(v)]
  opt = H.reference_options(options_mod, argv)
  with H.git_worktree(tmp_path):
    return trainer_mod.IkLagrangeDualTrainer(opt), opt


@pytest.mark.parametrize("name", CONFIG_NAMES)
def test_reference_process_batch_and_backward_on_the_dropins(name, tmp_path, cuda_device):
  from ikl import _cabi
  g = golden_npz("process_batch_%s.npz" % name)
  tr, opt = make_reference_trainer(name, tmp_path)
  assert list(g["constraint_names"]) == tr.constraint_names
  _, ospec = specs_for(name, float(g["good_slope"]), float(g["bad_slope"]))
  theta, q_des, obst = (torch.from_numpy(g[k]) for k in ("theta", "q_ee_desired", "dynamic_obstacles"))
  B = theta.shape[0]
  for tag in ("shipped", "ramp"):
    lam = g["lam_" + tag]
    _, abs_sums, dth64, kink, truth = oracle_truth(ospec, theta, q_des, obst, lam)
    tr.model = FixedTheta(theta).to(cuda_device)
    inputs = {"input_tensor": torch.zeros(B, 20), "q_ee_desired": q_des.clone(), "joint_theta": theta.clone()}   # host tensors, like the DataLoader's
    if tr.json_config["dynamic_constraints"]["random_obstacles"]:
      inputs["dynamic_obstacles"] = obst.clone()
    n0 = _cabi.launch_count()
    out = tr.process_batch(inputs, torch.from_numpy(lam).to(cuda_device))
    tr.model.zero_grad()
    out["lagrange_loss"].backward()
    launched = _cabi.launch_count() - n0
    # FK fwd + bwd, one fwd + bwd launch per penalty call: every piece of arithmetic of the path ran in libikl.so
    n_pen = (3 if tr.json_config["enforce_joint_limits"] else 0) + 3 * ospec.num_obstacles * (1 if tr.json_config["enforce_obstacles"] else 0)
    assert launched == 2 + 2 * n_pen, (launched, n_pen)
    for key in ("pred_joint_theta", "q_ee_actual", "constraint_violations", "lagrange_loss", "position_err_sq", "joint_angle_l2"):
      assert out[key].is_cuda, key
    viol = out["constraint_violations"].detach().cpu().numpy()
    assert np.all(np.abs(viol - g["viol_" + tag]) <= RTOL * abs_sums + 1e-30), (viol, g["viol_" + tag])
    close(float(out["lagrange_loss"]), g["loss_" + tag], scale=float(np.abs(lam * abs_sums).sum()))
    close(out["q_ee_actual"].detach().cpu().numpy(), g["q_ee_actual"])
    close(out["position_err_sq"].detach().cpu().numpy(), g["position_err_sq"])
    close(float(out["joint_angle_l2"]), g["joint_angle_l2"])
    dth = tr.model.theta.grad.detach().cpu().numpy()
    grad_close(dth, g["dtheta_" + tag], truth)
    grad_close(dth, dth64, truth)


@pytest.mark.parametrize("name", ("cfg_joint_limits_and_4obs_dynamic", "cfg_joint_limits_and_1obs_static"))
def test_reference_trainer_dual_trajectory_on_the_dropins(name, tmp_path, cuda_device):
  """6 dual epochs of the reference's train_with_relaxation + update_multipliers (trainer.py:251-286) on the GPU against
  the same loop run by the same file on the CPU with the reference's kinematics: lambda after every epoch, the final
  weights and an isolated dual step over 4 validation batches, <= 1e-5 of max(lambda)."""
  from models.simple_network import SimpleNetwork
  fx = torch.load(os.path.join(GOLDEN, "dual_trajectory_%s.pt" % name))
  o = fx["opt"]
  over = {k: o[k] for k in ("train_dataset_size", "val_dataset_size", "batch_size", "lagrange_iters", "train_iters", "optimizer_lr",
                             "multiplier_lr", "model_save_hz")}
  torch.manual_seed(1234)
  tr, opt = make_reference_trainer(name, tmp_path, **over)
  # the trainer built its datasets through the drop-in kinematics.dataset.IkDataset: they must be the reference's
  for split, ds in (("train", tr.train_dataset), ("val", tr.val_dataset)):
    assert torch.equal(ds.random_theta.cpu(), fx[split]["random_theta"])
    torch.testing.assert_close(ds.random_ee.cpu(), fx[split]["random_ee"], rtol=1e-5, atol=3e-6)
    if fx[split]["random_obstacles"] is not None:
      assert torch.equal(ds.random_obstacles.cpu(), fx[split]["random_obstacles"])
  # the golden run removed dropout (p = 0) so that it does not depend on a device's RNG stream
  tr.model = SimpleNetwork(20, 3, hidden_units=opt.hidden_units, depth=8, dropout=0.0).to(cuda_device)
  tr.model.load_state_dict(fx["init_state"])
  tr.optimizer = torch.optim.Adam(tr.model.parameters(), lr=opt.optimizer_lr, betas=(0.9, 0.999))
  hist = fx["lambda_history"]
  torch.testing.assert_close(tr.lambda_multipliers.cpu(), hist[0])
  for epoch in range(hist.shape[0] - 1):
    tr.train_with_relaxation(tr.lambda_multipliers)
    tr.lambda_multipliers = tr.update_multipliers(tr.lambda_multipliers)
    got, want = tr.lambda_multipliers.cpu().double(), hist[epoch + 1].double()
    assert float((got - want).abs().max()) <= 1e-5 * float(want.abs().max()), (epoch, got, want)
  for k, v in tr.model.state_dict().items():
    assert float((v.cpu() - fx["final_state"][k]).abs().max()) <= 2e-5 * max(1.0, float(fx["final_state"][k].abs().max())), k
  tr.model.load_state_dict(fx["final_state"])
  tr.val_loader = torch.utils.data.DataLoader(tr.val_dataset, 8, False, num_workers=0)
  lam_iso = tr.update_multipliers(hist[0].to(cuda_device))
  assert float((lam_iso.cpu() - fx["lambda_isolated_step"]).abs().max()) <= 1e-5 * float(fx["lambda_isolated_step"].abs().max())


def test_reference_main_loop_runs_end_to_end(tmp_path, cuda_device):
  """trainer.main() (trainer.py:325-341) for two dual epochs with the reference's defaults otherwise (dropout on, batch 8):
  validate() logs, checkpoints land in the reference's layout."""
  tr, opt = make_reference_trainer("cfg_joint_limits_and_4obs_dynamic", tmp_path, train_dataset_size=64, val_dataset_size=32,
                                   batch_size=8, lagrange_iters=2, train_iters=4, model_save_hz=1)
  tr.main()
  assert tr.lambda_multipliers.is_cuda and tr.lambda_multipliers.shape == (16,) and bool(torch.isfinite(tr.lambda_multipliers).all())
  assert os.path.exists(os.path.join(str(tmp_path), "golden", "models", "weights_1", "multipliers.pth"))
  assert os.path.exists(os.path.join(str(tmp_path), "golden", "constraint_names.txt"))
