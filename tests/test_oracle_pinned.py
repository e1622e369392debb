"""Pin the CPU oracle (oracle/ik_oracle.py) to the reference: its own known-answer tests, the golden
vectors produced by running the unmodified reference (tests/golden/make_golden.py), the shipped
multipliers' sizes, and an independent closed-form gradient.  No GPU needed."""
import glob
import json
import math
import os

import numpy as np
import pytest
import torch

from conftest import CONFIG_NAMES, GOLDEN, config_json, golden_npz
from oracle import ik_oracle as O

SHIPPED_K = {"cfg_no_constraints": 1, "cfg_joint_limits": 4, "cfg_joint_limits_and_1obs_static": 7,
             "cfg_joint_limits_and_4obs_static": 16, "cfg_joint_limits_and_4obs_dynamic": 16}


# --- the reference's own tests, restated against the oracle (test/test_training_utils.py:7-41) ------------------
def test_reference_known_answer_circle():
  jx = torch.Tensor([0.0, 0.5, 0.75, 0.5, -1.0])
  jy = torch.Tensor([0.0, 0.5, 0.5, 0.75, 0.5])
  got = O.circle_penalty(jx, jy, torch.Tensor([0.5]), torch.Tensor([0.5]), torch.Tensor([0.25]), inside_slope=1.0, outside_slope=0.1)
  want = torch.Tensor([-0.1 * (math.sqrt(0.5 ** 2 + 0.5 ** 2) - 0.25), 0.25, 0, 0, -0.1 * 1.25])
  torch.testing.assert_close(got, want)


def test_reference_known_answer_joint_limit():
  theta = torch.Tensor([-math.pi, math.pi, 0, 0.1])
  got = O.joint_limit_penalty(theta, torch.Tensor([-2 * math.pi / 3]), torch.Tensor([2 * math.pi / 3]), inside_slope=0.1, outside_slope=1.0)
  want = torch.Tensor([math.pi / 3, math.pi / 3, -0.1 * 2 * math.pi / 3, -0.1 * 2 * math.pi / 3 + 0.1 * 0.1])
  torch.testing.assert_close(got, want)


def test_negative_slopes_assert():
  with pytest.raises(AssertionError):
    O.circle_penalty(torch.zeros(1), torch.zeros(1), torch.zeros(1), torch.zeros(1), torch.ones(1), -1.0, 0.1)
  with pytest.raises(AssertionError):
    O.joint_limit_penalty(torch.zeros(1), torch.zeros(1), torch.ones(1), 0.1, -1.0)


# --- golden vectors from the reference itself -----------------------------------------------------------------
def test_fk_matches_reference_bitwise():
  g = golden_npz("fk.npz")
  theta = torch.from_numpy(g["theta"])
  x3, x2, x1 = O.forward_kinematics(theta)
  for got, key in ((x3, "x3"), (x2, "x2"), (x1, "x1")):
    assert np.array_equal(got.numpy(), g[key]), key
  y3, y2, y1 = O.forward_kinematics(theta.double())
  for got, key in ((y3, "x3_f64"), (y2, "x2_f64"), (y1, "x1_f64")):
    assert np.array_equal(got.numpy(), g[key]), key


def test_fk_hand_values():
  x3, x2, x1 = O.forward_kinematics(torch.zeros(1, 3))
  assert x1[0].tolist() == [1.0, 0.0, 0.0] and x2[0].tolist() == [2.0, 0.0, 0.0] and x3[0].tolist() == [3.0, 0.0, 0.0]
  x3, x2, x1 = O.forward_kinematics(torch.tensor([[math.pi / 2, -math.pi / 2, -math.pi / 2]]))
  np.testing.assert_allclose(x2[0].numpy(), [1.0, 1.0, 0.0], atol=1e-6)
  np.testing.assert_allclose(x3[0].numpy(), [1.0, 0.0, -math.pi / 2], atol=1e-6)


def test_fk_two_link_against_numpy_reference():
  g = golden_npz("fk.npz")
  th2 = torch.from_numpy(g["theta2"])
  ee = O.forward_kinematics(th2, (1.0, 1.0))[0]
  np.testing.assert_allclose(ee.numpy(), g["ee2_f64"], rtol=0, atol=1e-14)
  np.testing.assert_allclose(np.stack([O.fk_two_link_scalar(t) for t in g["theta2"]]), g["ee2_f64"], rtol=0, atol=1e-14)


def test_penalties_match_reference_bitwise_with_grads():
  g = golden_npz("penalties.npz")
  for i, (a_in, a_out) in enumerate(g["circle_slopes"]):
    for dt, tag in ((torch.float32, ""), (torch.float64, "_f64")):
      x = torch.from_numpy(g["jx"]).to(dt).requires_grad_(True)
      y = torch.from_numpy(g["jy"]).to(dt).requires_grad_(True)
      p = O.circle_penalty(x, y, torch.from_numpy(g["ox"]).to(dt), torch.from_numpy(g["oy"]).to(dt),
                           torch.from_numpy(g["radius"]).to(dt), float(a_in), float(a_out))
      p.sum().backward()
      assert np.array_equal(p.detach().numpy(), g["circle_%d%s" % (i, tag)])
      assert np.array_equal(x.grad.numpy(), g["circle_%d_djx%s" % (i, tag)])
      assert np.array_equal(y.grad.numpy(), g["circle_%d_djy%s" % (i, tag)])
  for li, (lo, hi) in enumerate(g["jl_limits"]):
    for si, (a_in, a_out) in enumerate(g["jl_slopes"]):
      for dt, tag in ((torch.float32, ""), (torch.float64, "_f64")):
        t = torch.from_numpy(g["jl_theta_%d" % li]).to(dt).requires_grad_(True)
        p = O.joint_limit_penalty(t, torch.Tensor([lo]).to(dt), torch.Tensor([hi]).to(dt), float(a_in), float(a_out))
        p.sum().backward()
        assert np.array_equal(p.detach().numpy(), g["jl_%d_%d%s" % (li, si, tag)])
        assert np.array_equal(t.grad.numpy(), g["jl_%d_%d_dtheta%s" % (li, si, tag)])


def test_circle_centre_reference_gradient_is_nan():
  """Documents the reference behaviour the CUDA path deliberately does not copy (DESIGN.md edge semantics)."""
  g = golden_npz("penalties.npz")
  assert g["circle_centre_value"][0] == np.float32(0.25)
  assert np.isnan(g["circle_centre_djx"][0])


@pytest.mark.parametrize("name", CONFIG_NAMES)
def test_lagrangian_matches_real_trainer(name):
  g = golden_npz("process_batch_%s.npz" % name)
  spec = O.spec_from_json(config_json(name), 3, float(g["good_slope"]), float(g["bad_slope"]))
  assert spec.num_constraints == SHIPPED_K[name] == len(g["lam_shipped"])
  assert O.constraint_names(spec) == [str(s) for s in g["constraint_names"]]
  theta = torch.from_numpy(g["theta"])
  q_des = torch.from_numpy(g["q_ee_desired"])
  obst = torch.from_numpy(g["dynamic_obstacles"])
  torch.set_num_threads(1)
  for tag in ("shipped", "ramp"):
    lam = torch.from_numpy(g["lam_" + tag])
    out = O.lagrangian_with_grad(theta, q_des, lam, spec, obst if spec.dynamic else None)
    assert np.array_equal(out["q_ee_actual"].numpy(), g["q_ee_actual"])
    assert np.array_equal(out["position_err_sq"].numpy(), g["position_err_sq"])
    np.testing.assert_allclose(out["constraint_violations"].numpy(), g["viol_" + tag], rtol=2e-6, atol=1e-6)
    np.testing.assert_allclose(out["lagrange_loss"].numpy(), g["loss_" + tag], rtol=2e-6)
    np.testing.assert_allclose(out["dtheta"].numpy(), g["dtheta_" + tag], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(out["joint_angle_l2"].numpy(), g["joint_angle_l2"], rtol=1e-6)


@pytest.mark.parametrize("name", CONFIG_NAMES)
def test_closed_form_gradient_agrees_with_autograd_fp64(name):
  g = golden_npz("process_batch_%s.npz" % name)
  spec = O.spec_from_json(config_json(name), 3, float(g["good_slope"]), float(g["bad_slope"]))
  theta = torch.from_numpy(g["theta"]).double()
  q_des = torch.from_numpy(g["q_ee_desired"]).double()
  obst = torch.from_numpy(g["dynamic_obstacles"]).double()
  lam = torch.from_numpy(g["lam_ramp"]).double()
  out = O.lagrangian_with_grad(theta, q_des, lam, spec, obst if spec.dynamic else None)
  viol, dth = O.lagrangian_grad_closed_form(theta.numpy(), q_des.numpy(), lam.numpy(), spec, obst.numpy() if spec.dynamic else None)
  np.testing.assert_allclose(viol, out["constraint_violations"].numpy(), rtol=1e-12, atol=1e-12)
  np.testing.assert_allclose(dth, out["dtheta"].numpy(), rtol=1e-11, atol=1e-12)


def test_survey_known_answer_vector():
  """SURVEY Appendix B.3 (fp32 reference run recorded at survey time)."""
  spec = O.spec_from_json(config_json("cfg_joint_limits_and_4obs_dynamic"), 3, 0.005, 1.0)
  theta = torch.tensor([[0.3, -0.5, 0.8], [2.5, -1.0, 0.2]])
  q_des = torch.tensor([[1.5, 1.0, 0.0], [-0.5, 2.0, 0.0]])
  ob = torch.zeros(2, 4, 4)
  ob[0, :, :2] = torch.tensor([[0.7, 0.7], [-0.7, 0.7], [0.7, -0.7], [-0.7, -0.7]])
  ob[1, :, :2] = torch.tensor([[1.0, 0.2], [1.9, 0.1], [-1.0, 0.5], [2.6, 0.9]])
  ob[:, :, 2] = 0.4
  lam = 1.0 + 0.25 * torch.arange(16, dtype=torch.float32)
  out = O.lagrangian_with_grad(theta, q_des, lam, spec, ob)
  np.testing.assert_allclose(out["q_ee_actual"][:, :2].numpy(), [[2.76073885, 0.66149342], [-0.85925096, 2.58763194]], rtol=1e-6)
  np.testing.assert_allclose(out["lagrange_loss"].item(), 1.45058775, rtol=1e-6)
  np.testing.assert_allclose(out["dtheta"].numpy(), [[-1.78350699, -1.11016190, -0.96378148], [5.12729836, 0.56544834, 0.26095459]], rtol=1e-6)
  want_g = [1.08921075, 0.39663309, -0.01344395, -0.01594395, -0.02143626, -0.01399034, -0.00761513, -0.03188014,
            -0.02464794, -0.01825395, -0.01881126, -0.00899369, 0.17495897, -0.03383942, -0.02677793, -0.02273057]
  np.testing.assert_allclose(out["constraint_violations"].numpy(), want_g, rtol=2e-6, atol=1e-7)


def test_dual_update_matches_real_trainer():
  for name in ("cfg_joint_limits_and_4obs_dynamic", "cfg_joint_limits_and_1obs_static"):
    fx = torch.load(os.path.join(GOLDEN, "dual_trajectory_%s.pt" % name))
    hist = fx["lambda_history"]
    assert hist.shape[1] == SHIPPED_K[name]
    assert bool((hist >= 0).all())
    # clamp/step semantics on a hand vector (trainer.py:280-284)
    lam = torch.tensor([1.0, 0.0, 0.5])
    viol = torch.tensor([[10.0, -5.0, -30.0], [10.0, -5.0, -30.0]])
    got = O.dual_update(lam, viol, 1e-2)
    torch.testing.assert_close(got, torch.tensor([1.2, 0.0, 0.0]))


def test_standard_configs_equal_reference_json_when_mounted():
  ref = "/root/reference/resources"
  if not os.path.isdir(ref):
    pytest.skip("reference not mounted (GPU box)")
  for p in glob.glob(os.path.join(ref, "cfg_*.json")):
    with open(p) as f:
      assert json.load(f) == config_json(os.path.basename(p)[:-5]), p


# --- the C restatement (oracle/ik_oracle.c) against the pinned PyTorch oracle -------------------------------------------
@pytest.mark.parametrize("name", CONFIG_NAMES)
def test_c_oracle_agrees_with_pinned_torch_oracle(name):
  from oracle import build_oracle as C
  g = golden_npz("process_batch_%s.npz" % name)
  spec = O.spec_from_json(config_json(name), 3, float(g["good_slope"]), float(g["bad_slope"]))
  theta, q_des, obst, lam = g["theta"], g["q_ee_desired"], g["dynamic_obstacles"], g["lam_ramp"]
  ref = O.lagrangian_with_grad(torch.from_numpy(theta).double(), torch.from_numpy(q_des).double(), torch.from_numpy(lam).double(), spec,
                               torch.from_numpy(obst).double() if spec.dynamic else None)
  sums, abs_sums, dth = C.lagrangian_f64(spec, theta, q_des, obst, lam)
  np.testing.assert_allclose(sums, ref["constraint_violations"].numpy(), rtol=1e-12, atol=1e-12)
  np.testing.assert_allclose(dth, ref["dtheta"].numpy(), rtol=1e-11, atol=1e-12)
  assert np.all(abs_sums >= np.abs(sums) - 1e-12)
  # float path against the reference's own fp32 trainer outputs
  sums32, dth32 = C.lagrangian_f32(spec, theta, q_des, obst, lam)
  np.testing.assert_allclose(sums32, g["viol_ramp"], rtol=1e-5, atol=1e-5)
  np.testing.assert_allclose(dth32, g["dtheta_ramp"], rtol=2e-5, atol=2e-5)


def test_c_oracle_large_batch_and_threads():
  from oracle import build_oracle as C
  from ikl import synthetic
  spec = O.spec_from_json(config_json("cfg_joint_limits_and_4obs_dynamic"), 3, 0.005, 1.0)
  theta, q_des, obst = synthetic.make_batch(1 << 16, seed=3)
  lam = np.linspace(0.5, 2.0, 16)
  sums, abs_sums, dth = C.lagrangian_f64(spec, theta.numpy(), q_des.numpy(), obst.numpy(), lam)
  ref = O.lagrangian(theta.double(), q_des.double(), torch.from_numpy(lam), spec, obst.double())
  np.testing.assert_allclose(sums, ref["constraint_violations"].numpy(), rtol=1e-10)
  assert C.max_threads() >= 1
