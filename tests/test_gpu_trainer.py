"""GPU tests of the callers either side of the hot path: dataset generation, the trainer mirror (fused vs op-by-op
process_batch), and lambda / weight trajectories against a run of the UNMODIFIED reference trainer
(tests/golden/dual_trajectory_*.pt, produced by tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, config_json

pytestmark = pytest.mark.gpu

FIXTURES = ("cfg_joint_limits_and_4obs_dynamic", "cfg_joint_limits_and_1obs_static")


def make_trainer(name, tmp_path, fx, extra=()):
  from ikl.configs import write_config
  from kinematics.dataset import IkDataset
  from models.simple_network import SimpleNetwork
  from scripts.options import Options
  from scripts.trainer import IkLagrangeDualTrainer
  cfg = config_json(name)
  os.makedirs(str(tmp_path), exist_ok=True)
  cfg_path = write_config(cfg, str(tmp_path / (name + ".json")))
  o = fx["opt"]
  argv = ["--model_name", "t", "--config_file", cfg_path, "--log_dir", str(tmp_path), "--num_workers", "0"]
  for k in ("batch_size", "lagrange_iters", "train_iters", "optimizer_lr", "multiplier_lr", "hidden_units", "initial_lambda",
            "penalty_good_slope", "penalty_bad_slope", "train_dataset_size", "val_dataset_size", "model_save_hz"):
    argv += ["--" + k, str(o[k])]
  opt = Options().parse(argv + list(extra))
  datasets = (IkDataset.from_tensors(cfg, fx["train"]), IkDataset.from_tensors(cfg, fx["val"]))
  tr = IkLagrangeDualTrainer(opt, datasets=datasets)
  # the fixture run removed dropout (p = 0) so that it does not depend on the device's RNG
  tr.model = SimpleNetwork(20, 3, hidden_units=opt.hidden_units, depth=8, dropout=0.0).to(tr.device)
  tr.model.load_state_dict(fx["init_state"])
  tr.optimizer = torch.optim.Adam(tr.model.parameters(), lr=opt.optimizer_lr, betas=(0.9, 0.999))
  return tr


@pytest.mark.parametrize("name", FIXTURES)
def test_dataset_generation_reproduces_the_reference(name, cuda_device):
  from kinematics.dataset import IkDataset
  fx = torch.load(os.path.join(GOLDEN, "dual_trajectory_%s.pt" % name))
  for split, n in (("train", 64), ("val", 32)):
    ds = IkDataset(n, 3, config_json(name))          # seed 0, like the reference trainer (trainer.py:150-153)
    ref = fx[split]
    assert len(ds) == n
    assert torch.equal(ds.tensor_template, ref["tensor_template"])
    assert torch.equal(ds.random_theta, ref["random_theta"]), "same random stream, same accept/reject decisions"
    torch.testing.assert_close(ds.random_ee, ref["random_ee"], rtol=1e-5, atol=3e-6)
    if ref["random_obstacles"] is None:
      assert ds.random_obstacles is None
    else:
      assert torch.equal(ds.random_obstacles, ref["random_obstacles"])
    item = ds[3]
    assert item["input_tensor"].shape == (20,) and item["q_ee_desired"].shape == (3,) and item["joint_theta"].shape == (3,)
    assert torch.equal(item["input_tensor"][:2], ds.random_ee[3, :2])
    if ds.random_obstacles is not None:
      assert item["dynamic_obstacles"].shape == (4, 4) and torch.equal(item["input_tensor"][4:], ds.random_obstacles[3].flatten())


def test_dataset_leaves_the_global_generator_where_the_reference_would(cuda_device):
  """Block-wise drawing must not consume more of torch's CPU stream than draw-by-draw rejection sampling."""
  from kinematics.dataset import IkDataset
  cfg = config_json("cfg_joint_limits_and_4obs_dynamic")
  ds = IkDataset(16, 3, cfg, seed=5)
  after = torch.rand(4)
  # replay by hand: N*J angles, then one 2-vector per accepted or rejected candidate
  torch.manual_seed(5)
  torch.empty(16, 3).uniform_(-2.094395, 2.094395)
  n_candidates = 0
  from kinematics.penalties import no_joint_collisions_circle
  from kinematics.forward_kinematics import ForwardKinematicsThreeLinkTorch
  tips = ForwardKinematicsThreeLinkTorch(ds.random_theta)
  for i in range(16):
    q = [t[i] for t in tips]
    for o in range(4):
      xy = torch.empty(2).uniform_(-1.2, 1.2)
      while not no_joint_collisions_circle(q, xy[0], xy[1], cfg["dynamic_constraints"]["random_obstacles_template"]):
        xy = torch.empty(2).uniform_(-1.2, 1.2)
      assert torch.equal(xy, ds.random_obstacles[i, o, :2])
  assert torch.equal(torch.rand(4), after)


@pytest.mark.parametrize("name", FIXTURES)
def test_fused_and_unfused_process_batch_agree(name, tmp_path, cuda_device):
  fx = torch.load(os.path.join(GOLDEN, "dual_trajectory_%s.pt" % name))
  tr = make_trainer(name, tmp_path, fx)
  inputs = next(iter(tr.train_loader))
  lam = torch.linspace(0.5, 2.0, len(tr.lambda_multipliers), device=tr.device)
  fused_out = tr.process_batch({k: v.clone() for k, v in inputs.items()}, lam)
  tr.model.zero_grad()
  fused_out["lagrange_loss"].backward()
  g_fused = [p.grad.clone() for p in tr.model.parameters()]
  plain = tr.process_batch_unfused({k: v.clone() for k, v in inputs.items()}, lam)
  tr.model.zero_grad()
  plain["lagrange_loss"].backward()
  assert set(plain.keys()) <= set(fused_out.keys())
  for key in ("pred_joint_theta", "q_ee_actual", "joint_angle_l2", "position_err_sq", "constraint_violations", "lagrange_loss"):
    a, b = fused_out[key].detach().cpu(), plain[key].detach().cpu()
    assert a.shape == b.shape, key
    scale = float(b.abs().max()) if key != "constraint_violations" else float(b.abs().sum())
    assert float((a - b).abs().max()) <= 1e-5 * max(scale, 1e-6), key
  for a, p in zip(g_fused, tr.model.parameters()):
    assert float((a - p.grad).abs().max()) <= 2e-5 * max(float(p.grad.abs().max()), 1e-6)


@pytest.mark.parametrize("name", FIXTURES)
def test_lambda_and_weight_trajectory_match_the_reference_trainer(name, tmp_path, cuda_device):
  """6 dual epochs (Adam step + dual-ascent step each) from the same initial weights and datasets as a run of the
  reference's own IkLagrangeDualTrainer on the CPU: lambda after every epoch within 1e-5 (relative to max lambda)."""
  fx = torch.load(os.path.join(GOLDEN, "dual_trajectory_%s.pt" % name))
  tr = make_trainer(name, tmp_path, fx)
  hist = fx["lambda_history"]
  torch.testing.assert_close(tr.lambda_multipliers.cpu(), hist[0])
  for epoch in range(hist.shape[0] - 1):
    tr.train_with_relaxation(tr.lambda_multipliers)
    tr.lambda_multipliers = tr.update_multipliers(tr.lambda_multipliers)
    got, want = tr.lambda_multipliers.cpu().double(), hist[epoch + 1].double()
    assert float((got - want).abs().max()) <= 1e-5 * float(want.abs().max()), (epoch, got, want)
  for k, v in tr.model.state_dict().items():
    assert float((v.cpu() - fx["final_state"][k]).abs().max()) <= 2e-5 * max(1.0, float(fx["final_state"][k].abs().max())), k
  # the isolated dual step over 4 validation batches of 8 (sum over batches, trainer.py:274-280)
  tr.model.load_state_dict(fx["final_state"])
  tr.val_loader = torch.utils.data.DataLoader(tr.val_dataset, 8, False, num_workers=0)
  lam_iso = tr.update_multipliers(hist[0].to(tr.device))
  assert float((lam_iso.cpu() - fx["lambda_isolated_step"]).abs().max()) <= 1e-5 * float(fx["lambda_isolated_step"].abs().max())


def test_trainer_main_runs_and_writes_the_reference_artifacts(tmp_path, cuda_device):
  """A miniature end-to-end run through main(): log files, checkpoints in the reference's layout, validate()."""
  from ikl.configs import write_config
  from scripts.options import Options
  from scripts.trainer import IkLagrangeDualTrainer
  name = "cfg_joint_limits_and_4obs_dynamic"
  cfg_path = write_config(config_json(name), str(tmp_path / "cfg.json"))
  opt = Options().parse(["--model_name", "mini", "--config_file", cfg_path, "--log_dir", str(tmp_path), "--num_workers", "0",
                         "--train_dataset_size", "64", "--val_dataset_size", "32", "--batch_size", "8", "--lagrange_iters", "3",
                         "--train_iters", "4", "--model_save_hz", "1", "--hidden_units", "16", "--initial_lambda", "1",
                         "--penalty_good_slope", "0.005", "--cache_save_path", str(tmp_path)])
  tr = IkLagrangeDualTrainer(opt)
  lam0 = tr.lambda_multipliers.clone()
  tr.main()
  log = tmp_path / "mini"
  assert (log / "opt.json").exists() and (log / "config.json").exists()
  names = (log / "constraint_names.txt").read_text().strip().splitlines()
  assert names[0] == "0,EE" and names[-1] == "15,DYN_OB4_J3" and len(names) == 16
  for e in (1, 2):
    for f in ("model.pth", "adam.pth", "multipliers.pth"):
      assert (log / "models" / ("weights_%d" % e) / f).exists()
  assert (tmp_path / "mini_train.pt").exists() and (tmp_path / "mini_val.pt").exists()      # dataset cache files
  assert tr.lambda_multipliers.shape == lam0.shape and bool((tr.lambda_multipliers >= 0).all())
  assert not torch.equal(tr.lambda_multipliers, lam0)
  saved = torch.load(log / "models" / "weights_2" / "multipliers.pth", map_location="cpu")
  assert torch.equal(saved, tr.lambda_multipliers.cpu())
  # a second trainer picks the cached datasets up and can resume weights + multipliers
  opt2 = Options().parse(["--model_name", "mini", "--config_file", cfg_path, "--log_dir", str(tmp_path), "--num_workers", "0",
                          "--train_dataset_size", "64", "--val_dataset_size", "32", "--hidden_units", "16",
                          "--cache_save_path", str(tmp_path), "--load_weights_folder", str(log / "models" / "weights_2"),
                          "--load_adam", "--load_multipliers"])
  tr2 = IkLagrangeDualTrainer(opt2)
  assert torch.equal(tr2.lambda_multipliers.cpu(), saved)
  for a, b in zip(tr.model.parameters(), tr2.model.parameters()):
    assert torch.equal(a, b)


# ---- device-side dataset generation and evaluation statistics -----------------------------------
@pytest.mark.parametrize("name", ("cfg_joint_limits_and_4obs_dynamic", "cfg_joint_limits_and_4obs_static", "cfg_joint_limits"))
def test_device_dataset_generation(name, cuda_device):
  from kinematics.dataset import IkDataset
  from kinematics.forward_kinematics import ForwardKinematicsThreeLinkTorch
  cfg = config_json(name)
  N = 1 << 16
  ds = IkDataset.generate_on_device(N, 3, cfg, seed=7)
  assert ds.random_theta.is_cuda and len(ds) == N
  lo, hi = cfg["static_constraints"]["joint_limits"]
  th = ds.random_theta
  assert float(th.min()) >= lo and float(th.max()) < hi
  tips = ForwardKinematicsThreeLinkTorch(th)
  assert torch.equal(ds.random_ee, tips[0]), "targets are the FK of the generated angles"
  # uniform angles (only the dynamic/unconstrained configs keep the raw uniform draw)
  if not cfg["static_constraints"]["obstacles"]:
    assert abs(float(th.mean())) < 0.02 and abs(float(th.var()) - (hi - lo) ** 2 / 12) < 0.03
  if ds.random_obstacles is not None:
    ob = ds.random_obstacles
    assert float(ob[:, :, :2].abs().max()) < 1.2 and torch.all(ob[:, :, 2] == 0.4) and torch.all(ob[:, :, 3] == 0)
    for o in range(4):
      for t in tips:
        d = torch.sqrt((t[:, 0] - ob[:, o, 0]) ** 2 + (t[:, 1] - ob[:, o, 1]) ** 2)
        assert float(d.min()) > 0.4, "no joint tip inside an obstacle"
      assert float(torch.sqrt(ob[:, o, 0] ** 2 + ob[:, o, 1] ** 2).min()) > 0.4, "arm base outside every obstacle"
  else:
    # static: the last obstacle checked is always cleared (earlier ones may not be: the reference's rule)
    obs = cfg["static_constraints"]["obstacles"]
    if obs:
      last = obs[-1]
      for t in tips:
        assert float(torch.sqrt((t[:, 0] - last["x"]) ** 2 + (t[:, 1] - last["y"]) ** 2).min()) > last["radius"]
  # deterministic, and a function of (seed, index) only
  again = IkDataset.generate_on_device(N, 3, cfg, seed=7)
  assert torch.equal(again.random_theta, ds.random_theta)
  prefix = IkDataset.generate_on_device(1000, 3, cfg, seed=7)
  assert torch.equal(prefix.random_theta, ds.random_theta[:1000])
  other = IkDataset.generate_on_device(1000, 3, cfg, seed=8)
  assert not torch.equal(other.random_theta, prefix.random_theta)
  item = ds[5]
  assert item["input_tensor"].shape == (20,) and torch.equal(item["input_tensor"][:2], ds.random_ee[5, :2])


def test_device_dataset_matches_host_generator_statistically(cuda_device):
  """Same procedure, different random stream: first and second moments of the accepted obstacle centres agree."""
  from kinematics.dataset import IkDataset
  cfg = config_json("cfg_joint_limits_and_4obs_dynamic")
  host = IkDataset(4000, 3, cfg, seed=1)
  dev = IkDataset.generate_on_device(1 << 17, 3, cfg, seed=1)
  h, d = host.random_obstacles[:, :, :2].reshape(-1, 2), dev.random_obstacles[:, :, :2].reshape(-1, 2).cpu()
  assert float((h.mean(0) - d.mean(0)).abs().max()) < 0.03
  assert float((h.var(0) - d.var(0)).abs().max()) < 0.03
  assert abs(float(h.norm(dim=1).mean()) - float(d.norm(dim=1).mean())) < 0.02


def test_eval_stats_match_per_sample_oracle(cuda_device, monkeypatch):
  from ikl import ConstraintSpec, synthetic, _cabi
  monkeypatch.delenv("IKL_FORCE_LAYOUT", raising=False)
  from ikl.device_data import eval_stats
  from oracle import ik_oracle as O
  for name in ("cfg_joint_limits_and_4obs_dynamic", "cfg_joint_limits_and_1obs_static", "cfg_no_constraints"):
    cfg = config_json(name)
    spec = ConstraintSpec.from_config(cfg, 3, 0.005, 1.0)
    ospec = O.spec_from_json(cfg, 3, 0.005, 1.0)
    theta, q_des, obst = synthetic.make_batch(5000, seed=31)
    out = O.lagrangian(theta.double(), q_des.double(), torch.ones(spec.num_constraints).double(), ospec, obst.double(), per_sample=True)
    terms = out["per_sample_terms"]
    K = spec.num_constraints
    want = (terms > 1e-4).sum(dim=0)
    names = O.constraint_names(ospec)
    jl = [i for i, n in enumerate(names) if "JL" in n]
    ob = [i for i, n in enumerate(names) if "OB" in n]
    want_err = torch.sqrt(out["position_err_sq"].sum(dim=1)).sum()
    th, qd, obd = theta.cuda(), q_des.cuda(), obst.cuda()
    planes = synthetic.to_planes(th, qd, obd, max(spec.num_obstacles, 1))
    big = torch.zeros(2 * len(theta), 3, device="cuda")
    big[::2] = th
    # the staged rows / planes pipelines (5000 = 9 full tiles + a ragged one) and the generic strided kernel
    for args, kernel in (((th, qd, obd), "rows_tma"), (planes, "planes_tma"), ((big[::2], qd, obd), "strided")):
      counts, err = eval_stats(spec, args[0], args[1], args[2] if spec.dynamic else None)
      assert kernel in _cabi.last_kernel_name(), (_cabi.last_kernel_name(), kernel)
      counts = counts.cpu()
      assert int((counts[:K] - want).abs().max()) <= 2, (counts[:K], want)        # fp32 vs fp64 right at the threshold
      assert int(counts[K:16].abs().max() if K < 16 else 0) == 0
      if jl:
        assert abs(int(counts[16]) - int((terms[:, jl] > 0).any(dim=1).sum())) <= 2
      if ob:
        assert abs(int(counts[17]) - int((terms[:, ob] > 0).any(dim=1).sum())) <= 2
      assert abs(float(err) - float(want_err)) <= 1e-5 * float(want_err)


def test_val_batch_size_gives_the_same_dual_step(tmp_path, cuda_device):
  """--val_batch_size: the dual step over the validation set in ONE launch instead of one per 8 examples -- the sums are
  additive, so lambda moves the same way (dropout is off in the fixtures; the reference's train-mode validation would draw
  different masks)."""
  name = "cfg_joint_limits_and_4obs_dynamic"
  fx = torch.load(os.path.join(GOLDEN, "dual_trajectory_%s.pt" % name))
  small = make_trainer(name, tmp_path / "small", fx, extra=("--val_batch_size", "8"))
  n_val = len(small.val_dataset)
  big = make_trainer(name, tmp_path / "big", fx, extra=("--val_batch_size", str(n_val)))
  assert len(big.val_loader) == 1 and len(small.val_loader) > 1
  lam_s = small.update_multipliers(small.lambda_multipliers)
  lam_b = big.update_multipliers(big.lambda_multipliers)
  assert float((lam_s - lam_b).abs().max()) <= 1e-6 * float(lam_s.abs().max())
  assert not torch.equal(lam_s, small.lambda_multipliers)


def test_cuda_graph_train_step_matches_eager(tmp_path, cuda_device):
  """--cuda_graph replays forward + fused Lagrangian + backward + Adam as one graph; same weights and lambda as eager."""
  name = "cfg_joint_limits_and_4obs_dynamic"
  fx = torch.load(os.path.join(GOLDEN, "dual_trajectory_%s.pt" % name))
  eager = make_trainer(name, tmp_path / "eager", fx)
  (tmp_path / "graph").mkdir()
  graph = make_trainer(name, tmp_path / "graph", fx, extra=("--cuda_graph",))
  graph.optimizer = torch.optim.Adam(graph.model.parameters(), lr=graph.opt.optimizer_lr, betas=(0.9, 0.999), capturable=True)
  assert graph.use_graph
  for tr in (eager, graph):
    for _ in range(4):
      tr.train_with_relaxation(tr.lambda_multipliers)
      tr.lambda_multipliers = tr.update_multipliers(tr.lambda_multipliers)
  assert graph._graph_step is not None
  assert float((eager.lambda_multipliers - graph.lambda_multipliers).abs().max()) <= 1e-5 * float(eager.lambda_multipliers.abs().max())
  for a, b in zip(eager.model.parameters(), graph.model.parameters()):
    assert float((a - b).detach().abs().max()) <= 2e-5 * max(1.0, float(a.detach().abs().max()))
  hist = fx["lambda_history"]
  assert float((graph.lambda_multipliers.cpu() - hist[4]).abs().max()) <= 2e-5 * float(hist[4].abs().max())


def test_evaluation_script_matches_per_example_loop(tmp_path, cuda_device):
  """scripts/evaluation/evaluate_ik_solutions.py against the reference's procedure (visualize_ik_solutions.py:53-119):
  one example at a time through process_batch, thresholds on the host."""
  from scripts.evaluation.evaluate_ik_solutions import evaluate
  name = "cfg_joint_limits_and_4obs_dynamic"
  fx = torch.load(os.path.join(GOLDEN, "dual_trajectory_%s.pt" % name))
  tr = make_trainer(name, tmp_path, fx)
  tr.model.load_state_dict(fx["final_state"])
  stats = evaluate(tr, batch_size=16)
  tr.model.eval()
  pos, viol = [], []
  with torch.no_grad():
    for i in range(len(tr.val_dataset)):
      inputs = {k: v.unsqueeze(0) for k, v in tr.val_dataset[i].items()}
      out = tr.process_batch_unfused(inputs, tr.lambda_multipliers)
      pos.append(torch.sqrt(out["position_err_sq"].sum()).unsqueeze(0))
      viol.append(out["constraint_violations"].unsqueeze(0))
  viol = torch.cat(viol).cpu()
  want_counts = (viol > 1e-4).sum(dim=0)
  names = tr.constraint_names
  jl = [i for i, n in enumerate(names) if "JL" in n]
  ob = [i for i, n in enumerate(names) if "OB" in n]
  assert stats["n"] == 32
  assert stats["num_violations"] == want_counts.tolist()
  assert stats["any_joint_limit"] == int(((viol[:, jl] > 0).sum(dim=1) > 0).sum())
  assert stats["any_obstacle"] == int(((viol[:, ob] > 0).sum(dim=1) > 0).sum())
  assert abs(stats["position_err"] - float(torch.cat(pos).mean())) <= 1e-5 * float(torch.cat(pos).mean())


def test_trainer_with_device_resident_data(tmp_path, cuda_device):
  """--device_data: datasets generated on the GPU, batches assembled there; main() runs with a large batch and lambda moves."""
  from ikl.configs import write_config
  from scripts.options import Options
  from scripts.trainer import IkLagrangeDualTrainer
  cfg_path = write_config(config_json("cfg_joint_limits_and_4obs_dynamic"), str(tmp_path / "cfg.json"))
  opt = Options().parse(["--model_name", "dev", "--config_file", cfg_path, "--log_dir", str(tmp_path), "--device_data",
                         "--train_dataset_size", "65536", "--val_dataset_size", "16384", "--batch_size", "8192", "--lagrange_iters", "2",
                         "--train_iters", "8", "--hidden_units", "32", "--initial_lambda", "1", "--penalty_good_slope", "0.005",
                         "--model_save_hz", "1000"])
  tr = IkLagrangeDualTrainer(opt)
  assert tr.train_dataset.random_theta.is_cuda and len(tr.train_loader) == 8 and len(tr.val_loader) == 2
  # the device-resident path is planes end to end: the bulk-async kernel runs inside the real trainer step, and it
  # computes the same loss and weight gradients as the row-major path on the same batch
  from ikl import _cabi
  tr.model.eval()
  batch = next(iter(tr.train_loader))
  assert batch["q_ee_desired"].stride(0) == 1 and batch["dynamic_obstacles"].stride(0) == 1
  out_p = tr.process_batch(dict(batch), tr.lambda_multipliers)
  assert "planes_tma" in _cabi.last_kernel_name() and out_p["pred_joint_theta"].stride(0) == 1
  tr.model.zero_grad()
  out_p["lagrange_loss"].backward()
  g_p = [p.grad.clone() for p in tr.model.parameters()]
  rows = {"input_tensor": batch["input_tensor"], "q_ee_desired": batch["q_ee_desired"].contiguous(),
          "dynamic_obstacles": torch.cat([batch["dynamic_obstacles"], torch.zeros(8192, 4, 1, device=tr.device)], dim=2).contiguous()}
  out_r = tr.process_batch(rows, tr.lambda_multipliers)
  assert "<rows" in _cabi.last_kernel_name()
  tr.model.zero_grad()
  out_r["lagrange_loss"].backward()
  lp, lr = float(out_p["lagrange_loss"].detach()), float(out_r["lagrange_loss"].detach())
  assert abs(lp - lr) <= 1e-5 * abs(lr)
  for a, p in zip(g_p, tr.model.parameters()):
    assert float((a - p.grad).abs().max()) <= 2e-5 * max(float(p.grad.abs().max()), 1e-6)
  assert torch.equal(out_p["q_ee_actual"], out_r["q_ee_actual"])
  tr.model.train()
  tr.model.zero_grad()
  lam0 = tr.lambda_multipliers.clone()
  w0 = [p.detach().clone() for p in tr.model.parameters()]
  tr.main()
  assert not torch.equal(tr.lambda_multipliers, lam0) and bool((tr.lambda_multipliers >= 0).all())
  assert any(not torch.equal(a, b) for a, b in zip(w0, tr.model.parameters()))
  assert all(torch.isfinite(p).all() for p in tr.model.parameters())


def test_graph_capture_keeps_a_loaded_optimizer_state(cuda_device):
  """--cuda_graph --load_adam: capturing the step warms up on dummy data, which must not disturb the Adam moments and the
  step count that were there before (a resumed run), nor the weights; a fresh optimizer must come out at step 0."""
  from ikl import ConstraintSpec, synthetic
  from ikl.graph_step import GraphedTrainStep
  from ikl import fused
  from models.simple_network import SimpleNetwork
  spec = ConstraintSpec.from_config(config_json("cfg_joint_limits_and_4obs_dynamic"), 3, 0.005, 1.0)
  B = 64
  _, q_des, obst = synthetic.make_batch(B, seed=5, device="cuda:0")
  x = torch.randn(B, 20, device=cuda_device)
  inputs = {"input_tensor": x, "q_ee_desired": q_des, "dynamic_obstacles": obst}
  lam = torch.linspace(0.5, 2.0, 16, device=cuda_device)
  torch.manual_seed(3)
  model = SimpleNetwork(20, 3, hidden_units=16, depth=8, dropout=0.0).to(cuda_device)
  opt = torch.optim.Adam(model.parameters(), lr=1e-3, capturable=True)

  def eager_step(m, o):
    viol, loss = fused.FusedIkLagrangian.apply(m(x), q_des, obst, lam, spec)
    o.zero_grad()
    loss.backward()
    o.step()
  for _ in range(3):                                     # "the run before the checkpoint": builds non-trivial Adam state
    eager_step(model, opt)
  import copy
  twin = copy.deepcopy(model)
  twin_opt = torch.optim.Adam(twin.parameters(), lr=1e-3, capturable=True)
  twin_opt.load_state_dict(copy.deepcopy(opt.state_dict()))
  before = {k: v.clone() for k, v in model.state_dict().items()}
  before_opt = [{k: v.clone() for k, v in opt.state[p].items()} for p in model.parameters()]
  g = GraphedTrainStep(model, opt, spec, B, cuda_device)
  for k, v in model.state_dict().items():
    assert torch.equal(v, before[k]), k
  for p, st in zip(model.parameters(), before_opt):
    for k, v in st.items():
      assert torch.equal(opt.state[p][k], v), k
    assert float(opt.state[p]["step"]) == 3.0
  g.set_multipliers(lam)
  for _ in range(2):
    g.step(inputs)
    eager_step(twin, twin_opt)
  for a, b in zip(model.parameters(), twin.parameters()):
    assert float((a - b).detach().abs().max()) <= 2e-6 * max(1.0, float(b.detach().abs().max()))
  # a fresh optimizer: capture leaves it at step 0 with zero moments
  m2 = SimpleNetwork(20, 3, hidden_units=16, depth=8, dropout=0.0).to(cuda_device)
  o2 = torch.optim.Adam(m2.parameters(), lr=1e-3, capturable=True)
  GraphedTrainStep(m2, o2, spec, B, cuda_device)
  for p in m2.parameters():
    assert float(o2.state[p]["step"]) == 0.0 and float(o2.state[p]["exp_avg"].abs().max()) == 0.0


def test_device_data_with_cuda_graph_on_a_dynamic_config(tmp_path, cuda_device):
  """--device_data --cuda_graph: the loaders hand (B,4,3) obstacle PLANES; the graph's static buffers are planes too and the
  TMA-staged planes kernel is the one captured.  Same weights as the eager --device_data run from the same state."""
  from ikl.configs import write_config
  from ikl import _cabi
  from models.simple_network import SimpleNetwork
  from scripts.options import Options
  from scripts.trainer import IkLagrangeDualTrainer
  cfg_path = write_config(config_json("cfg_joint_limits_and_4obs_dynamic"), str(tmp_path / "cfg.json"))
  trainers = []
  for tag, extra in (("eager", []), ("graph", ["--cuda_graph"])):
    opt = Options().parse(["--model_name", tag, "--config_file", cfg_path, "--log_dir", str(tmp_path), "--device_data",
                           "--train_dataset_size", "16384", "--val_dataset_size", "4096", "--batch_size", "4096", "--lagrange_iters", "2",
                           "--train_iters", "4", "--hidden_units", "32", "--initial_lambda", "1", "--penalty_good_slope", "0.005",
                           "--model_save_hz", "1000"] + extra)
    tr = IkLagrangeDualTrainer(opt)
    torch.manual_seed(11)
    tr.model = SimpleNetwork(20, 3, hidden_units=32, depth=8, dropout=0.0).to(tr.device)
    tr.optimizer = torch.optim.Adam(tr.model.parameters(), lr=opt.optimizer_lr, betas=(0.9, 0.999), capturable=tr.use_graph)
    trainers.append(tr)
  eager, graph = trainers
  graph.model.load_state_dict(eager.model.state_dict())
  assert graph.use_graph and not eager.use_graph
  for tr in trainers:
    for _ in range(2):
      tr.train_with_relaxation(tr.lambda_multipliers)
      tr.lambda_multipliers = tr.update_multipliers(tr.lambda_multipliers)
  assert graph._graph_step is not None and graph._graph_step.planes
  assert graph._graph_step.obst.stride(0) == 1 and graph._graph_step.q_des.stride(0) == 1
  assert float((eager.lambda_multipliers - graph.lambda_multipliers).abs().max()) <= 1e-5 * float(eager.lambda_multipliers.abs().max())
  for a, b in zip(eager.model.parameters(), graph.model.parameters()):
    assert float((a - b).detach().abs().max()) <= 2e-5 * max(1.0, float(a.detach().abs().max()))


def test_unit_upstream_skips_the_scaling_pass_and_is_checked(cuda_device, monkeypatch):
  from ikl import ConstraintSpec, synthetic, fused
  spec = ConstraintSpec.from_config(config_json("cfg_joint_limits_and_4obs_dynamic"), 3, 0.005, 1.0)
  theta, q_des, obst = synthetic.make_batch(1024, seed=9, device="cuda:0")
  lam = torch.linspace(0.5, 2.0, 16, device=cuda_device)
  grads = []
  for unit in (False, True):
    th = theta.clone().requires_grad_(True)
    viol, loss = fused.FusedIkLagrangian.apply(th, q_des, obst, lam, spec, None, unit)
    loss.backward()
    grads.append(th.grad.clone())
  assert torch.equal(grads[0], grads[1])
  # the promise is checkable: a scaled loss with unit_upstream=True trips the (opt-in) assertion
  monkeypatch.setattr(fused, "_CHECK_UNIT_UPSTREAM", True)
  th = theta.clone().requires_grad_(True)
  viol, loss = fused.FusedIkLagrangian.apply(th, q_des, obst, lam, spec, None, True)
  with pytest.raises(AssertionError):
    (2.0 * loss).backward()


def test_training_steps_use_the_gradient_only_kernel(tmp_path, cuda_device):
  """train_with_relaxation consumes only lagrange_loss.backward() (reference trainer.py:262-265), so its launches are the
  gradient-only kernel unless --train_sums: the weights after a relaxation epoch must be bit-identical either way, the
  autograd function must hand back NaN placeholders (never stale sums) and refuse gradients it cannot serve, and the dual
  step / validation must still see real sums."""
  from ikl import ConstraintSpec, synthetic, fused, _cabi
  name = "cfg_joint_limits_and_4obs_dynamic"
  fx = torch.load(os.path.join(GOLDEN, "dual_trajectory_%s.pt" % name))
  from ikl.configs import write_config
  from kinematics.dataset import IkDataset
  from models.simple_network import SimpleNetwork
  from scripts.options import Options
  from scripts.trainer import IkLagrangeDualTrainer
  cfg = config_json(name)
  os.makedirs(str(tmp_path), exist_ok=True)
  cfg_path = write_config(cfg, str(tmp_path / (name + ".json")))
  datasets = (IkDataset(2048, 3, cfg), IkDataset(1024, 3, cfg))
  finals = []
  for extra in ((), ("--train_sums",)):
    argv = ["--model_name", "t", "--config_file", cfg_path, "--log_dir", str(tmp_path), "--num_workers", "0", "--batch_size", "1024",
            "--train_iters", "2", "--hidden_units", "32", "--train_dataset_size", "2048", "--val_dataset_size", "1024"]
    tr = IkLagrangeDualTrainer(Options().parse(argv + list(extra)), datasets=datasets)
    torch.manual_seed(0)
    tr.model = SimpleNetwork(20, 3, hidden_units=32, depth=8, dropout=0.0).to(tr.device)
    tr.optimizer = torch.optim.Adam(tr.model.parameters(), lr=1e-3, betas=(0.9, 0.999))
    lam = tr.lambda_multipliers.clone() + 0.5
    tr.train_with_relaxation(lam)
    kernel = _cabi.last_kernel_name()
    assert ("grad_" in kernel) == (not extra), kernel
    lam2 = tr.update_multipliers(lam)
    assert torch.isfinite(lam2).all() and ",fwd," in _cabi.last_kernel_name()
    finals.append([p.detach().clone() for p in tr.model.parameters()] + [lam2])
  for a, b in zip(*finals):
    assert torch.equal(a, b)
  # the autograd function itself
  spec = ConstraintSpec.from_config(config_json(name), 3, 0.005, 1.0)
  theta, q_des, obst = synthetic.make_batch(1024, seed=9, device="cuda:0")
  lam = torch.linspace(0.5, 2.0, 16, device=cuda_device)
  th = theta.clone().requires_grad_(True)
  viol, loss = fused.FusedIkLagrangian.apply(th, q_des, obst, lam, spec, None, True, True)
  assert torch.isnan(viol).all() and torch.isnan(loss)
  loss.backward()
  th2 = theta.clone().requires_grad_(True)
  fused.FusedIkLagrangian.apply(th2, q_des, obst, lam, spec, None, True)[1].backward()
  assert torch.equal(th.grad, th2.grad)
  with pytest.raises(ValueError):
    fused.FusedIkLagrangian.apply(th, q_des, obst, lam, spec, None, False, True)
  th3 = theta.clone().requires_grad_(True)
  viol, loss = fused.FusedIkLagrangian.apply(th3, q_des, obst, lam, spec, None, True, True)
  with pytest.raises(RuntimeError):
    viol.sum().backward()
