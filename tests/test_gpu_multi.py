"""Two-GPU tests (skipped on a single-GPU box): the trainer under torch.distributed/NCCL keeps lambda and the weights
identical on every rank and matches a single-GPU run on the same global batches."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import FUSED_OVERLAY, GOLDEN, PKG, ROOT, config_json

pytestmark = pytest.mark.gpu
NAME = "cfg_joint_limits_and_4obs_dynamic"


def _build(tmpdir, tag):
  from test_gpu_trainer import make_trainer
  import pathlib
  fx = torch.load(os.path.join(GOLDEN, "dual_trajectory_%s.pt" % NAME))
  d = pathlib.Path(tmpdir) / tag
  d.mkdir(parents=True, exist_ok=True)
  return make_trainer(NAME, d, fx), fx


def _worker(rank, world, port, tmpdir):
  os.environ["MASTER_ADDR"] = "127.0.0.1"
  os.environ["MASTER_PORT"] = str(port)
  for p in (ROOT, FUSED_OVERLAY, PKG, os.path.join(ROOT, "tests")):
    if p not in sys.path:
      sys.path.insert(0, p)
  torch.cuda.set_device(rank)
  dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
  try:
    tr, fx = _build(tmpdir, "rank%d" % rank)
    assert tr.world == world and tr.train_loader.batch_sampler.batch_size == 64
    for _ in range(3):
      tr.train_with_relaxation(tr.lambda_multipliers)
      tr.lambda_multipliers = tr.update_multipliers(tr.lambda_multipliers)
    lam = tr.lambda_multipliers.clone()
    flat = torch.cat([p.detach().reshape(-1) for p in tr.model.parameters()])
    lams = [torch.zeros_like(lam) for _ in range(world)]
    flats = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(lams, lam)
    dist.all_gather(flats, flat)
    assert all(torch.equal(l, lams[0]) for l in lams), "lambda must be bit-identical on every rank"
    assert all(torch.equal(f, flats[0]) for f in flats), "replicas must stay in lock-step"
    want = fx["lambda_history"][3].to(lam.device)       # the reference's single-process run on the same global batches
    assert float((lam - want).abs().max()) <= 2e-5 * float(want.abs().max())
    open(os.path.join(tmpdir, "ok_%d" % rank), "w").write("ok")
  finally:
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_gpu_trainer_matches_single_process_reference(tmp_path):
  s = socket.socket()
  s.bind(("127.0.0.1", 0))
  port = s.getsockname()[1]
  s.close()
  mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
  assert (tmp_path / "ok_0").exists() and (tmp_path / "ok_1").exists()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_one_process_driving_two_devices():
  """The staged kernels opt in to > 48 KiB of dynamic shared memory, a per-device function attribute: one process that
  launches them on cuda:0 and then on cuda:1 must get the same numbers on both (and no launch error on the second)."""
  from ikl import ConstraintSpec, fused, synthetic, _cabi
  from ikl.device_data import eval_stats
  spec = ConstraintSpec.from_config(config_json(NAME), 3, 0.005, 1.0)
  B = 4096
  theta, q_des, obst = synthetic.make_batch(B, seed=3)
  lam = [0.5 + 0.1 * i for i in range(16)]
  results = []
  for d in (0, 1):
    dev = torch.device("cuda", d)
    th, qd, ob = theta.to(dev), q_des.to(dev), obst.to(dev)
    out = []
    for args in ((th, qd, ob), synthetic.to_planes(th, qd, ob)):
      v, l, g = fused.lagrangian_forward_backward(spec, *args, lam)
      assert "_tma," in _cabi.last_kernel_name()
      counts, err = eval_stats(spec, *args)
      torch.cuda.synchronize(dev)
      out += [v.cpu(), g.contiguous().cpu(), counts.cpu(), err.cpu()]
    results.append(out)
  for a, b in zip(*results):
    assert torch.equal(a, b)


def _exchange_worker(rank, world, port, tmpdir):
  os.environ["MASTER_ADDR"] = "127.0.0.1"
  os.environ["MASTER_PORT"] = str(port)
  for p in (ROOT, FUSED_OVERLAY, PKG, os.path.join(ROOT, "tests")):
    if p not in sys.path:
      sys.path.insert(0, p)
  torch.cuda.set_device(rank)
  dev = torch.device("cuda", rank)
  dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
  try:
    from ikl import ConstraintSpec, fused, synthetic, _cabi
    from ikl.exchange import PeerExchange
    ex = PeerExchange(device=dev)
    for name in ("cfg_joint_limits_and_4obs_dynamic", "cfg_joint_limits_and_1obs_static", "cfg_no_constraints"):
      spec = ConstraintSpec.from_config(config_json(name), 3, 0.005, 1.0)
      K = spec.num_constraints
      lam = [0.5 + 0.1 * i for i in range(K)]
      for B in (4096 + 512 * rank, 8):                      # ragged across ranks; tiny (strided kernel) too
        theta, q_des, obst = synthetic.make_batch(B, seed=17 + rank, device=str(dev))
        for layout in ("rows", "planes"):
          args = synthetic.to_planes(theta, q_des, obst, max(spec.num_obstacles, 1)) if layout == "planes" else (theta, q_des, obst)
          ob = args[2] if spec.dynamic else None
          local, lloss, g_local = fused.lagrangian_forward_backward(spec, args[0], args[1], ob, lam)
          accum = torch.zeros(K, dtype=torch.float64, device=dev)
          for rep in range(3):                                # consecutive launches alternate the mailbox parity
            glob, gloss, g = fused.lagrangian_forward_backward(spec, args[0], args[1], ob, lam, viol_accum=accum, exchange=ex)
          assert torch.equal(g, g_local), "d/dtheta of the shard does not depend on the exchange"
          parts = [torch.zeros(K, dtype=torch.float64, device=dev) for _ in range(world)]
          dist.all_gather(parts, local.double())
          want = torch.stack(parts).sum(0)
          scale = torch.stack(parts).abs().sum(0).clamp(min=1e-30)
          assert float(((glob.double() - want).abs() / scale).max()) <= 1e-6, (name, B, layout)
          assert float(((accum / 3 - want).abs() / scale).max()) <= 1e-6
          both = [torch.zeros_like(glob) for _ in range(world)]
          dist.all_gather(both, glob)
          assert torch.equal(both[0], both[1]), "global sums must be bit-identical on every rank"
          want_loss = float((torch.tensor(lam, dtype=torch.float64, device=dev) * want).sum())
          assert abs(float(gloss) - want_loss) <= 1e-5 * float((torch.tensor(lam, dtype=torch.float64, device=dev) * scale).sum())
          fwd = fused.lagrangian_forward(spec, args[0], args[1], ob, lam, exchange=ex)
          assert torch.equal(fwd["constraint_violations"], glob), "forward-only and fwd+bwd launches exchange the same sums"
    # the launch that closes a dual step (fused.DualStep): local running sums, ONE exchange per epoch, lambda moved in the kernel
    for name in ("cfg_joint_limits_and_4obs_dynamic", "cfg_joint_limits_and_4obs_static", "cfg_joint_limits"):
      spec = ConstraintSpec.from_config(config_json(name), 3, 0.005, 1.0)
      K = spec.num_constraints
      lam = [0.5 + 0.1 * i for i in range(K)]
      lam_dev = torch.tensor(lam, device=dev)
      for layout in ("rows", "planes"):
        for epoch in range(2):                               # interleaves with the per-launch exchange: one shared sequence counter
          local_total = torch.zeros(K, dtype=torch.float64, device=dev)
          accum = torch.zeros(K, dtype=torch.float64, device=dev)
          sizes = (4096 + 512 * rank, 1024, 2048 * (1 - rank))        # the closing batch is EMPTY on rank 1
          for i, B in enumerate(sizes):
            theta, q_des, obst = synthetic.make_batch(max(B, 1), seed=50 + 10 * epoch + i + 100 * rank, device=str(dev))
            theta, q_des, obst = theta[:B], q_des[:B], obst[:B]
            args = synthetic.to_planes(theta, q_des, obst, max(spec.num_obstacles, 1)) if (layout == "planes" and B > 0) else (theta, q_des, obst)
            ob = args[2] if spec.dynamic else None
            fused.lagrangian_forward(spec, args[0], args[1], ob, lam, viol_accum=local_total)
            close = fused.DualStep(lam_dev, 0.01, exchange=ex) if i == len(sizes) - 1 else None
            out = fused.lagrangian_forward(spec, args[0], args[1], ob, lam, viol_accum=accum, close_epoch=close)
          parts = [torch.zeros(K, dtype=torch.float64, device=dev) for _ in range(world)]
          dist.all_gather(parts, local_total)
          want = parts[0] + parts[1]                        # the kernel adds the ranks' running sums in rank order, in float64
          assert torch.equal(accum, want), (name, layout, epoch, rank)
          assert torch.equal(close.lam_out, fused.dual_update(lam_dev, want, 0.01))
          lams = [torch.zeros_like(lam_dev) for _ in range(world)]
          dist.all_gather(lams, close.lam_out)
          assert torch.equal(lams[0], lams[1]), "lambda must be bit-identical on every rank"
          fused.lagrangian_forward(spec, args[0], args[1], ob, lam, exchange=ex)       # a per-launch exchange in between
    # CUDA-graph replay: the sequence counter lives on the device, so replays pair up across ranks
    spec = ConstraintSpec.from_config(config_json(NAME), 3, 0.005, 1.0)
    lam = [1.0] * 16
    theta, q_des, obst = synthetic.make_batch(8192, seed=40 + rank, device=str(dev))
    out = torch.empty(17, device=dev)
    dth = torch.empty_like(theta)
    side = torch.cuda.Stream(dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
      fused.lagrangian_forward_backward(spec, theta, q_des, obst, lam, dtheta=dth, out=out, exchange=ex)
      graph = torch.cuda.CUDAGraph()
      with torch.cuda.graph(graph, stream=side):
        for _ in range(5):
          fused.lagrangian_forward_backward(spec, theta, q_des, obst, lam, dtheta=dth, out=out, exchange=ex)
    torch.cuda.current_stream(dev).wait_stream(side)
    first = out.clone()
    for _ in range(4):
      graph.replay()
    torch.cuda.synchronize(dev)
    assert torch.equal(out, first)
    both = [torch.zeros_like(out) for _ in range(world)]
    dist.all_gather(both, out)
    assert torch.equal(both[0], both[1])
    open(os.path.join(tmpdir, "ex_ok_%d" % rank), "w").write("ok")
  finally:
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_in_kernel_peer_exchange_of_the_sums(tmp_path):
  """The fused kernel all-reduces the K sums itself over NVLink peer memory (include/ikl.h IklExchange): global sums are
  bit-identical on both ranks, equal the float64 sum of the local sums, and the shard gradient is untouched."""
  s = socket.socket()
  s.bind(("127.0.0.1", 0))
  port = s.getsockname()[1]
  s.close()
  mp.spawn(_exchange_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
  assert (tmp_path / "ex_ok_0").exists() and (tmp_path / "ex_ok_1").exists()


def _graph_worker(rank, world, port, tmpdir):
  os.environ["MASTER_ADDR"] = "127.0.0.1"
  os.environ["MASTER_PORT"] = str(port)
  for p in (ROOT, FUSED_OVERLAY, PKG, os.path.join(ROOT, "tests")):
    if p not in sys.path:
      sys.path.insert(0, p)
  torch.cuda.set_device(rank)
  dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
  try:
    import pathlib
    from test_gpu_trainer import make_trainer
    fx = torch.load(os.path.join(GOLDEN, "dual_trajectory_%s.pt" % NAME))
    trainers = []
    for tag, extra in (("eager", ()), ("graph", ("--cuda_graph",))):
      d = pathlib.Path(tmpdir) / ("%s_rank%d" % (tag, rank))
      d.mkdir(parents=True, exist_ok=True)
      tr = make_trainer(NAME, d, fx, extra=extra)
      if tag == "graph":
        tr.optimizer = torch.optim.Adam(tr.model.parameters(), lr=tr.opt.optimizer_lr, betas=(0.9, 0.999), capturable=True)
        assert tr.use_graph and tr.world == world and tr._peer_allreduce is not None
      trainers.append(tr)
    for tr in trainers:
      for _ in range(3):
        tr.train_with_relaxation(tr.lambda_multipliers)
        tr.lambda_multipliers = tr.update_multipliers(tr.lambda_multipliers)
    eager, graph = trainers
    assert graph._graph_step is not None and graph._graph_step.batch_size == 64 // world
    assert float((eager.lambda_multipliers - graph.lambda_multipliers).abs().max()) <= 1e-5 * float(eager.lambda_multipliers.abs().max())
    for a, b in zip(eager.model.parameters(), graph.model.parameters()):
      assert float((a - b).detach().abs().max()) <= 2e-5 * max(1.0, float(a.detach().abs().max()))
    flat = torch.cat([graph.lambda_multipliers.reshape(-1)] + [p.detach().reshape(-1) for p in graph.model.parameters()])
    flats = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(flats, flat)
    assert all(torch.equal(f, flats[0]) for f in flats), "graph replicas must stay in lock-step"
    want = fx["lambda_history"][3].to(flat.device)
    assert float((graph.lambda_multipliers - want).abs().max()) <= 2e-5 * float(want.abs().max())
    open(os.path.join(tmpdir, "graph_ok_%d" % rank), "w").write("ok")
  finally:
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.timeout(240)
def test_two_gpu_cuda_graph_step_captures_the_gradient_all_reduce(tmp_path):
  """--cuda_graph under torch.distributed: forward, gradient-only fused Lagrangian, backward, the all-reduce of the gradients
  (one kernel over NVLink peer memory, ikl.exchange.PeerAllReduce -- the captured NCCL all-reduce hung on this pool's
  two-GPU box) and Adam are ONE graph replay per step on every rank; same lambda / weights as the eager two-rank run and
  as the reference's single-process run on the same global batches."""
  s = socket.socket()
  s.bind(("127.0.0.1", 0))
  port = s.getsockname()[1]
  s.close()
  mp.spawn(_graph_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
  assert (tmp_path / "graph_ok_0").exists() and (tmp_path / "graph_ok_1").exists()


def _allreduce_worker(rank, world, port, tmpdir):
  os.environ["MASTER_ADDR"] = "127.0.0.1"
  os.environ["MASTER_PORT"] = str(port)
  for p in (ROOT, PKG, os.path.join(ROOT, "tests")):
    if p not in sys.path:
      sys.path.insert(0, p)
  dev = torch.device("cuda", rank)
  torch.cuda.set_device(dev)
  dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
  try:
    from ikl.exchange import PeerAllReduce
    ar = PeerAllReduce(100000, device=dev)
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    for n in (1, 255, 2563, 83203, 100000, 0):
      for rep in range(3):                       # consecutive launches alternate the mailbox parity
        x = torch.randn(n, generator=g, device=dev) * (10.0 ** (rep - 1))
        if n == 0:
          assert ar.all_reduce_(x).numel() == 0
          continue
        parts = [torch.zeros_like(x) for _ in range(world)]
        dist.all_gather(parts, x)
        want = parts[0].clone()
        for q in parts[1:]:
          want += q                              # rank order, fp32: what the kernel adds
        got = ar.all_reduce_(x.clone())
        assert torch.equal(got, want), (n, rep)
        both = [torch.zeros_like(got) for _ in range(world)]
        dist.all_gather(both, got)
        assert all(torch.equal(b, both[0]) for b in both), "the sum must be bit-identical on every rank"
    with pytest.raises(ValueError):
      ar.all_reduce_(torch.zeros(100001, device=dev))
    # CUDA-graph capture: the launch counter lives in the mailbox and is advanced by the kernel, so replays pair up
    x = torch.randn(2563, generator=g, device=dev)
    buf = x.clone()
    side = torch.cuda.Stream(dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
      ar.all_reduce_(buf)
      buf.copy_(x)
      graph = torch.cuda.CUDAGraph()
      with torch.cuda.graph(graph, stream=side):
        for _ in range(3):
          ar.all_reduce_(buf)                    # buf <- world^3 * (sum of the ranks' x) over one replay
    torch.cuda.current_stream(dev).wait_stream(side)
    parts = [torch.zeros_like(x) for _ in range(world)]
    dist.all_gather(parts, x)
    want = parts[0].clone()
    for q in parts[1:]:
      want += q
    for _ in range(4):
      buf.copy_(x)
      graph.replay()
      torch.cuda.synchronize(dev)
      torch.testing.assert_close(buf, want * float(world ** 2), rtol=1e-6, atol=1e-6)
    # interleaved with NCCL collectives and eager launches afterwards
    y = ar.all_reduce_(torch.ones(7, device=dev) * (rank + 1))
    assert torch.equal(y, torch.full((7,), float(sum(range(1, world + 1))), device=dev))
    open(os.path.join(tmpdir, "ar_ok_%d" % rank), "w").write("ok")
  finally:
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.timeout(240)
def test_peer_memory_all_reduce_of_a_flat_vector(tmp_path):
  """ikl_allreduce_sum_f32: the gradient all-reduce of the data-parallel step as one kernel over NVLink peer memory -- equal
  to the rank-order fp32 sum, bit-identical on every rank, ragged sizes, consecutive launches, CUDA-graph replays."""
  s = socket.socket()
  s.bind(("127.0.0.1", 0))
  port = s.getsockname()[1]
  s.close()
  mp.spawn(_allreduce_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
  assert (tmp_path / "ar_ok_0").exists() and (tmp_path / "ar_ok_1").exists()
