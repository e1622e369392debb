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
    assert tr.world == world and tr.train_loader.batch_size == 64 // world
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
