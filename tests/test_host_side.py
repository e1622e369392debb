"""CPU tests of the host-side mirror: options, checkpoint helpers, network, constraint layout, multi-process plumbing
(gloo, world_size 2).  Nothing here launches a kernel."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import FUSED_OVERLAY, GOLDEN, PKG, ROOT, config_json
from oracle import ik_oracle as O

REF = "/root/reference"


def test_options_have_the_reference_flags_and_defaults():
  from scripts.options import Options
  opt = Options().parse(["--model_name", "m", "--config_file", "c.json"])
  want = dict(lagrange_iters=1000, train_iters=500, batch_size=8, train_dataset_size=32000, val_dataset_size=4000,
              multiplier_lr=1e-4, optimizer_lr=1e-4, initial_lambda=40, num_links=3, num_workers=4, model_save_hz=200,
              hidden_units=40, penalty_bad_slope=1.0, penalty_good_slope=0.1, load_adam=False, load_weights_folder=None,
              cache_save_path=None, show_groundtruth_theta=False, no_plot=False)
  for k, v in want.items():
    assert getattr(opt, k) == v, k
  with pytest.raises(SystemExit):
    Options().parse(["--num_links", "5"])


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference not mounted (GPU box)")
def test_options_match_reference_parser_when_mounted():
  import importlib.util
  from scripts.options import Options
  spec = importlib.util.spec_from_file_location("ref_options", os.path.join(REF, "scripts", "options.py"))
  mod = importlib.util.module_from_spec(spec)
  spec.loader.exec_module(mod)         # its `from utils.paths import top_folder` resolves to our utils.paths
  ref_actions = {a.dest: a for a in mod.Options().parser._actions if a.dest != "help"}
  our_actions = {a.dest: a for a in Options().parser._actions if a.dest != "help"}
  for dest, a in ref_actions.items():
    assert dest in our_actions, dest
    if dest != "log_dir":
      assert our_actions[dest].default == a.default, dest
    assert our_actions[dest].type == a.type, dest


def test_simple_network_matches_reference_architecture():
  from models.simple_network import SimpleNetwork
  from utils.training_utils import count_parameters
  net = SimpleNetwork(20, 3, hidden_units=100, depth=8, dropout=0.05)
  assert count_parameters(net) == 83203
  keys = list(net.state_dict().keys())
  assert keys[:2] == ["input.weight", "input.bias"] and "hidden.7.0.weight" in keys and keys[-1] == "output.bias"
  assert net(torch.zeros(5, 20)).shape == (5, 3)
  fx = torch.load(os.path.join(GOLDEN, "dual_trajectory_cfg_joint_limits_and_4obs_dynamic.pt"))
  small = SimpleNetwork(20, 3, hidden_units=16, depth=8, dropout=0.0)
  small.load_state_dict(fx["init_state"])                   # state_dict written by the reference's own SimpleNetwork


def test_checkpoint_helpers_roundtrip(tmp_path):
  from models.simple_network import SimpleNetwork
  from utils import training_utils as tu
  net = SimpleNetwork(20, 3, hidden_units=8, depth=2)
  adam = torch.optim.Adam(net.parameters(), lr=1e-3)
  net(torch.randn(4, 20)).sum().backward()
  adam.step()
  mp_, ap_ = tu.save_model(net, adam, str(tmp_path), 7)
  assert mp_.endswith(os.path.join("weights_7", "model.pth")) and ap_.endswith(os.path.join("weights_7", "adam.pth"))
  lam = torch.tensor([1.0, 0.0, 2.5])
  lp = tu.save_multipliers(lam, str(tmp_path), 7)
  assert lp.endswith(os.path.join("weights_7", "multipliers.pth"))
  torch.testing.assert_close(torch.load(lp), lam)            # raw tensor, the reference's format
  torch.testing.assert_close(tu.load_multipliers(str(tmp_path), 7), lam)
  net2 = SimpleNetwork(20, 3, hidden_units=8, depth=2)
  adam2 = torch.optim.Adam(net2.parameters(), lr=1e-3)
  tu.load_model(net2, adam2, mp_, ap_)
  for a, b in zip(net.parameters(), net2.parameters()):
    assert torch.equal(a, b)
  assert tu.save_model(net, None, str(tmp_path), 8)[1] is None
  assert tu.get_best_system_device() in ("cuda", "mps", "cpu")


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference not mounted (GPU box)")
def test_shipped_multipliers_load_and_fit_the_constraint_layout():
  from ikl import ConstraintSpec
  from utils.training_utils import load_multipliers
  table = {"cfg_no_constraints": "ik_threelink_position/models/weights_600", "cfg_joint_limits": "ik_threelink_joint_limits/models/weights_200",
           "cfg_joint_limits_and_1obs_static": "ik_threelink_obs1/models/weights_950",
           "cfg_joint_limits_and_4obs_static": "ik_threelink_obs4/models/weights_950",
           "cfg_joint_limits_and_4obs_dynamic": "ik_threelink_obs4_dyn/models/weights_650"}
  for name, folder in table.items():
    lam = load_multipliers(os.path.join(REF, "resources/pretrained_models", folder))      # saved on a CUDA device
    spec = ConstraintSpec.from_config(config_json(name), 3)
    assert lam.dtype == torch.float32 and lam.numel() == spec.num_constraints == len(spec.constraint_names)


# ---- multi-process (gloo, world_size 2) ---------------------------------------------------------------------------
def _worker(rank, world, port, tmpdir):
  os.environ["MASTER_ADDR"] = "127.0.0.1"
  os.environ["MASTER_PORT"] = str(port)
  for p in (ROOT, FUSED_OVERLAY, PKG):
    if p not in sys.path:
      sys.path.insert(0, p)
  dist.init_process_group("gloo", rank=rank, world_size=world)
  try:
    from ikl import dist as D, synthetic
    from ikl.configs import standard_config
    from models.simple_network import SimpleNetwork
    assert D.is_distributed() and D.rank() == rank and D.world_size() == world
    # replicas start identical
    torch.manual_seed(100 + rank)
    net = SimpleNetwork(20, 3, hidden_units=8, depth=2, dropout=0.0)
    D.broadcast_parameters(net)
    flat = torch.cat([p.detach().reshape(-1) for p in net.parameters()])
    gathered = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    assert all(torch.equal(g, gathered[0]) for g in gathered)
    # the dual-learning exchange: batch sharded contiguously, one flat all-reduce of [grads | K sums]
    cfg = standard_config("cfg_joint_limits_and_4obs_dynamic")
    ospec = O.spec_from_json(cfg, 3, 0.005, 1.0)
    B = 103                                             # ragged on purpose
    theta_in = torch.randn(B, 20, generator=torch.Generator().manual_seed(5))
    _, q_des, obst = synthetic.make_batch(B, seed=3)
    lam = torch.linspace(0.5, 2.0, 16)
    lo, hi = D.shard_bounds(B)
    assert (lo, hi) == ((0, 52) if rank == 0 else (52, 103))
    out = O.lagrangian(net(theta_in[lo:hi]), q_des[lo:hi], lam, ospec, obst[lo:hi])
    net.zero_grad()
    out["lagrange_loss"].backward()
    viol = out["constraint_violations"].detach().clone()
    class NeverUsed(object):                         # a peer-memory all-reduce serves CUDA float32 buffers only: CPU tensors go to the group
      max_elems = 1 << 30

      def all_reduce_(self, flat):
        raise AssertionError("the peer all-reduce must not be handed a CPU tensor")
    D.all_reduce_gradients(net, extra=viol, peer=NeverUsed())
    # single-process truth on the whole batch
    ref_net = SimpleNetwork(20, 3, hidden_units=8, depth=2, dropout=0.0)
    ref_net.load_state_dict(net.state_dict())
    ref = O.lagrangian(ref_net(theta_in), q_des, lam, ospec, obst)
    ref["lagrange_loss"].backward()
    for p, q in zip(net.parameters(), ref_net.parameters()):
      torch.testing.assert_close(p.grad, q.grad, rtol=2e-5, atol=1e-5)
    torch.testing.assert_close(viol, ref["constraint_violations"].detach(), rtol=1e-5, atol=1e-5)
    # dual step: float64 running sums, all-reduce, identical lambda everywhere
    total = out["constraint_violations"].detach().double()
    D.all_reduce_sum(total)
    lam_new = O.dual_update(lam, total.float().unsqueeze(0), 1e-3)
    both = [torch.zeros_like(lam_new) for _ in range(world)]
    dist.all_gather(both, lam_new)
    assert torch.equal(both[0], both[1]), "lambda must be bit-identical on every rank"
    torch.testing.assert_close(lam_new, O.dual_update(lam, ref["constraint_violations"].detach().unsqueeze(0), 1e-3), rtol=1e-5, atol=1e-6)
    # loaders: the ranks' shares of an epoch partition the dataset
    class Items(torch.utils.data.Dataset):          # dict items, like IkDataset.__getitem__ (kinematics/dataset.py:143-155)
      def __len__(self):
        return 41                                    # NOT divisible by the world size or the batch size

      def __getitem__(self, i):
        return {"idx": torch.tensor(i), "row": torch.full((3,), float(i))}
    loader = D.make_loader(Items(), 8, 0)
    batches = list(loader)
    assert len(batches) == 6 and batches[0]["idx"].shape == (4,) and batches[0]["row"].shape == (4, 3)
    assert batches[-1]["idx"].shape == ((1,) if rank == 0 else (0,)) and batches[-1]["row"].shape[1:] == (3,)    # ragged tail: 1 + 0
    seen = torch.cat([b["idx"] for b in batches])
    padded = torch.full((21,), -1, dtype=seen.dtype)
    padded[:len(seen)] = seen
    allseen = [torch.zeros_like(padded) for _ in range(world)]
    dist.all_gather(allseen, padded)
    got = torch.cat(allseen)
    assert sorted(got[got >= 0].tolist()) == list(range(41)), "every sample exactly once: nothing padded, nothing dropped"
    with open(os.path.join(tmpdir, "ok_%d" % rank), "w") as f:
      f.write("ok")
  finally:
    dist.destroy_process_group()


def test_two_rank_gloo_exchange(tmp_path):
  import socket
  s = socket.socket()
  s.bind(("127.0.0.1", 0))
  port = s.getsockname()[1]
  s.close()
  mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
  assert os.path.exists(tmp_path / "ok_0") and os.path.exists(tmp_path / "ok_1")


def test_single_process_dist_helpers_are_noops():
  from ikl import dist as D
  assert not D.is_distributed() and D.rank() == 0 and D.world_size() == 1
  assert D.shard_bounds(10) == (0, 10) and D.shard_bounds(10, 1, 3) == (4, 7) and D.shard_bounds(10, 2, 3) == (7, 10)
  t = torch.ones(3)
  assert D.all_reduce_sum(t) is t


def test_device_batch_loader_covers_the_dataset_and_shards_by_rank():
  """CPU tensors exercise the same indexing logic the GPU path uses."""
  from ikl.configs import standard_config
  from ikl.device_loader import DeviceBatchLoader
  from kinematics.dataset import IkDataset
  cfg = standard_config("cfg_joint_limits_and_4obs_dynamic")
  n = 50
  tensors = {"random_theta": torch.arange(n * 3, dtype=torch.float32).reshape(n, 3), "random_ee": torch.rand(n, 3),
             "random_obstacles": torch.rand(n, 4, 4), "tensor_template": torch.cat([torch.zeros(2), torch.tensor([-2.0, 2.0]), torch.zeros(16)])}
  ds = IkDataset.from_tensors(cfg, tensors)
  loader = DeviceBatchLoader(ds, 16, shuffle=True, rank=0, world=1, seed=3)
  assert len(loader) == 4
  seen = []
  for b in loader:
    assert set(b.keys()) == {"input_tensor", "joint_theta", "q_ee_desired", "dynamic_obstacles"}
    assert b["input_tensor"].shape[1] == 20 and torch.equal(b["input_tensor"][:, :2], b["q_ee_desired"][:, :2])
    assert torch.equal(b["input_tensor"][:, 4:], b["dynamic_obstacles"].reshape(-1, 16))
    assert torch.equal(b["input_tensor"][:, 2:4], torch.tensor([-2.0, 2.0]).expand(len(b["joint_theta"]), 2))
    seen.append(b["joint_theta"][:, 0] / 3)
  assert sorted(torch.cat(seen).long().tolist()) == list(range(n))
  # item-by-item equality with the reference-style __getitem__
  first = next(iter(DeviceBatchLoader(ds, 8, shuffle=False, rank=0, world=1)))
  for i in range(8):
    item = ds[i]
    for k in item:
      assert torch.equal(first[k][i], item[k]), k
  # two ranks split every global batch without overlap
  parts = [list(DeviceBatchLoader(ds, 16, shuffle=True, rank=r, world=2, seed=5)) for r in (0, 1)]
  for b0, b1 in zip(*parts):
    ids = torch.cat([b0["joint_theta"][:, 0], b1["joint_theta"][:, 0]]) / 3
    assert len(set(ids.long().tolist())) == len(ids)
  all_ids = torch.cat([b["joint_theta"][:, 0] for p in parts for b in p]) / 3
  assert sorted(all_ids.long().tolist()) == list(range(n))


# ---- sharding without padding, lazy outputs ---------------------------------------------------------------------------
@pytest.mark.parametrize("n,bs,world", [(4001, 8, 3), (32, 8, 2), (10, 4, 4), (7, 16, 2), (4000, 8, 8)])
def test_sharded_batch_sampler_counts_every_sample_exactly_once(n, bs, world):
  """The dual step adds the constraint sums over the WHOLE validation set (trainer.py:274-280): under any world size every
  sample must be seen exactly once (DistributedSampler(drop_last=False) repeats samples when n % world != 0), and every
  rank must see the same number of batches so that the collectives of a step pair up."""
  from ikl.dist import ShardedBatchSampler
  per_rank = [list(ShardedBatchSampler(n, bs, r, world, shuffle=True, seed=3)) for r in range(world)]
  assert len({len(b) for b in per_rank}) == 1 and len(per_rank[0]) == (n + bs - 1) // bs
  seen = sorted(i for batches in per_rank for b in batches for i in b)
  assert seen == list(range(n))
  for g in range(len(per_rank[0])):                                  # a global batch is dealt round-robin
    sizes = [len(per_rank[r][g]) for r in range(world)]
    assert max(sizes) - min(sizes) <= 1 and sum(sizes) == min(bs, n - g * bs)
  s0 = ShardedBatchSampler(n, bs, 0, world, shuffle=True, seed=3)
  first = list(s0)
  s0.set_epoch(1)
  assert list(s0) != first or n <= 1


def test_loader_collates_an_empty_share():
  from ikl.dist import _CollateWithEmpty
  ds = [{"a": torch.zeros(3), "b": torch.zeros(4, 4)} for _ in range(5)]
  c = _CollateWithEmpty(ds)
  out = c([])
  assert out["a"].shape == (0, 3) and out["b"].shape == (0, 4, 4)
  assert c(ds[:2])["b"].shape == (2, 4, 4)


def test_batch_outputs_behaves_like_the_reference_dict():
  """process_batch's outputs are a plain dict in the reference (trainer.py:163): get / items / values / iteration / len /
  dict() must all expose the lazily materialised entries."""
  from scripts.trainer import BatchOutputs
  calls = []

  def side(out):
    calls.append("side")
    dict.__setitem__(out, "q_ee_actual", 1)
    dict.__setitem__(out, "position_err_sq", 2)
    out._lazy.pop("q_ee_actual", None)
    out._lazy.pop("position_err_sq", None)

  def l2(out):
    calls.append("l2")
    dict.__setitem__(out, "joint_angle_l2", 3)

  def fresh():
    o = BatchOutputs({"q_ee_actual": side, "position_err_sq": side, "joint_angle_l2": l2})
    o["lagrange_loss"] = 0
    return o
  o = fresh()
  assert len(o) == 4 and "q_ee_actual" in o and set(o.keys()) == {"lagrange_loss", "q_ee_actual", "position_err_sq", "joint_angle_l2"}
  assert calls == []                                      # nothing materialised yet
  assert o.get("q_ee_actual") == 1 and o.get("position_err_sq") == 2 and calls == ["side"]      # one launch fills both
  assert o.get("nope", 7) == 7
  assert dict(fresh()) == {"lagrange_loss": 0, "q_ee_actual": 1, "position_err_sq": 2, "joint_angle_l2": 3}
  assert sorted(fresh().values()) == [0, 1, 2, 3]
  assert sorted(k for k, _ in fresh().items()) == sorted(fresh().keys()) == sorted(iter(fresh()))
  assert fresh()["joint_angle_l2"] == 3
  with pytest.raises(KeyError):
    fresh()["missing"]


def test_dropin_wrappers_describe_the_trainers_operands_without_copies():
  """Host logic of ikl/ops.py (no launch): the operands scripts/trainer.py passes -- column views of (B,3) / (B,4,4) tensors,
  expand()-ed limits, the expanded gradient of `.sum()` -- are handed to the kernels as (pointer, element stride) with no
  copy; anything else is broadcast (and copied only when it is not expressible as one stride)."""
  from ikl import ops
  B = 64
  q = torch.randn(B, 3)
  ob = torch.rand(B, 4, 4)
  lim = torch.Tensor([-2.0, 2.0]).unsqueeze(0).expand(B, -1)
  shape = torch.Size([B])
  for t, stride in ((q[:, 1], 3), (ob[:, 2, 0], 16), (lim[:, 0], 0), (torch.ones(()).expand(B), 0), (torch.arange(2.0 * B)[::2], 2)):
    v, s = ops._flat_view(t, shape)
    assert s == stride and v.data_ptr() == t.data_ptr(), (t.shape, t.stride(), s)
  assert ops._flat_view(torch.ones(1), torch.Size([1]))[1] == 0
  # a scalar against the batch, and a 2-D broadcast that no single stride describes
  v, s = ops._flat_view(torch.tensor(0.4), shape)
  assert s == 0
  v, s = ops._flat_view(q[:, 0:1], torch.Size([B, 4]))
  assert s == 1 and v.is_contiguous() and v.shape == (B, 4)
  v, s = ops._flat_view(ob[0, :, 2], torch.Size([B, 4]))
  assert s == 1 and v.is_contiguous()
  # _prep: the trainer's call (float32, one device, one shape) passes through untouched; other inputs are converted / refused
  cpu = torch.device("cpu")
  same, shp = ops._prep(cpu, q[:, 0], q[:, 1], ob[:, 0, 0])
  assert shp == shape and [a.data_ptr() for a in same] == [q[:, 0].data_ptr(), q[:, 1].data_ptr(), ob[:, 0, 0].data_ptr()]
  mixed, shp = ops._prep(cpu, q[:, 0:1], 0.5, ob[0, :, 2])
  assert shp == (B, 4) and mixed[1].dtype == torch.float32
  with pytest.raises(TypeError):
    ops._prep(cpu, q[:, 0], q[:, 1].double())
