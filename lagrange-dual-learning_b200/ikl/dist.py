"""Multi-GPU plumbing of the dual-learning step (one process per GPU, torch.distributed; NCCL over NVLink on a B200
box, gloo in the CPU tests).

The reference has no distributed code; the path shards trivially because every constraint value is a SUM over samples
(scripts/trainer.py:195,204,239) and the loss gradient is therefore a sum of per-shard gradients:

  train step   every rank runs the fused kernel on its shard of the batch, back-propagates through its replica of the
               network, then ONE all-reduce(sum) of a flat fp32 buffer of all parameter gradients (optionally followed by
               the K violation sums: `extra=`, used by bench.py's train step; the trainer does not read the sums while
               training) makes the gradients global; identical Adam steps keep the replicas in lock-step.
  dual step    ranks accumulate the K sums over their validation shard (float64 on device), one all-reduce(sum) of K
               doubles, then the same lambda <- max(0, lambda + lr * sum) everywhere.  NCCL returns bit-identical
               results on all ranks, so lambda cannot drift.

No rescaling by world size anywhere: sums, not means.
"""
import torch
import torch.distributed as dist
from torch.utils.data import DataLoader, Sampler
from torch.utils.data._utils.collate import default_collate


def is_distributed():
  return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def rank():
  return dist.get_rank() if is_distributed() else 0


def world_size():
  return dist.get_world_size() if is_distributed() else 1


def broadcast_parameters(model, src=0):
  """Start every replica from rank 0's initial weights."""
  if not is_distributed():
    return
  with torch.no_grad():
    flat = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    dist.broadcast(flat, src)
    offset = 0
    for p in model.parameters():
      n = p.numel()
      p.copy_(flat[offset:offset + n].view_as(p))
      offset += n


def all_reduce_gradients(model, extra=None, peer=None):
  """Sum parameter gradients (and, if given, the 1-D fp32 tensor `extra`, e.g. the K violation sums) over ranks in
  one flat collective.  `extra` is updated in place.  No-op on a single process.
  `peer` (ikl.exchange.PeerAllReduce): the sum runs as one kernel over NVLink peer memory instead of an NCCL all-reduce --
  no NCCL launch, capturable in the step's CUDA graph, bit-identical on every rank."""
  if not is_distributed():
    return
  grads = [p.grad for p in model.parameters() if p.grad is not None]
  pieces = [g.reshape(-1) for g in grads]
  if extra is not None:
    pieces.append(extra.detach().reshape(-1).to(pieces[0].dtype if pieces else extra.dtype))
  if not pieces:
    return
  flat = torch.cat(pieces)
  if peer is not None and flat.is_cuda and flat.dtype == torch.float32 and flat.numel() <= peer.max_elems:
    peer.all_reduce_(flat)
  else:
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
  offset = 0
  for g in grads:
    n = g.numel()
    g.copy_(flat[offset:offset + n].view_as(g))
    offset += n
  if extra is not None:
    with torch.no_grad():
      extra.copy_(flat[offset:offset + extra.numel()].view_as(extra))


def all_reduce_sum(t):
  """In-place sum over ranks (the K running violation sums of the dual step)."""
  if is_distributed():
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
  return t


def shard_bounds(n, r=None, w=None):
  """[start, end) of rank r's contiguous shard of n items (ragged tails go to the low ranks)."""
  r = rank() if r is None else r
  w = world_size() if w is None else w
  base, rem = divmod(n, w)
  start = r * base + min(r, rem)
  return start, start + base + (1 if r < rem else 0)


class ShardedBatchSampler(Sampler):
  """Deals every GLOBAL batch of a shuffled epoch over the ranks: global batch g is perm[g*bs:(g+1)*bs] (the same
  permutation on every rank: seeded by (seed, epoch)) and rank r takes elements r, r+world, ... of it.

  Unlike DistributedSampler(drop_last=False) nothing is padded, so no sample is counted twice in the dual step's sums
  (update_multipliers adds over the whole validation set, scripts/trainer.py:274-280) whatever len(dataset) % world is;
  unlike drop_last=True nothing is dropped.  Every rank sees the same NUMBER of batches (ceil(n / bs)) -- the collectives
  of a step must match up -- and a rank's share of a ragged last batch may be smaller than the others' or empty."""

  def __init__(self, n, batch_size, rank, world, shuffle=True, seed=0):
    self.n, self.batch_size, self.rank, self.world = int(n), int(batch_size), int(rank), int(world)
    self.shuffle, self.seed, self.epoch = shuffle, int(seed), 0

  def set_epoch(self, epoch):
    self.epoch = int(epoch)

  def __len__(self):
    return (self.n + self.batch_size - 1) // self.batch_size

  def __iter__(self):
    if self.shuffle:
      g = torch.Generator()
      g.manual_seed(self.seed + self.epoch)
      order = torch.randperm(self.n, generator=g).tolist()
    else:
      order = list(range(self.n))
    for s in range(0, self.n, self.batch_size):
      yield order[s:s + self.batch_size][self.rank::self.world]


class _CollateWithEmpty(object):
  """default_collate, except that an empty share of a ragged last batch becomes zero-row tensors of the item's shapes."""

  def __init__(self, dataset):
    self.dataset = dataset

  def __call__(self, items):
    if items:
      return default_collate(items)
    proto = default_collate([self.dataset[0]])
    return {k: v[:0] for k, v in proto.items()}


def make_loader(dataset, batch_size, num_workers, shuffle=True):
  """The reference's DataLoader(dataset, batch_size, shuffle=True) (trainer.py:155-156); under torch.distributed each rank
  draws its share of every global batch (batch_size is the GLOBAL batch size) through ShardedBatchSampler."""
  if not is_distributed():
    return DataLoader(dataset, batch_size, shuffle, num_workers=num_workers)
  sampler = ShardedBatchSampler(len(dataset), batch_size, rank(), world_size(), shuffle=shuffle)
  return DataLoader(dataset, batch_sampler=sampler, num_workers=num_workers, collate_fn=_CollateWithEmpty(dataset))
