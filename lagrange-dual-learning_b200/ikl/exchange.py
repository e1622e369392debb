"""Peer-memory mailboxes for the in-kernel all-reduce of the K constraint sums (include/ikl.h: IklExchange).

One process per GPU (torch.distributed, NCCL).  Each rank allocates ikl_exchange_bytes() of SYMMETRIC memory
(torch.distributed._symmetric_memory: cuMem allocations mapped into every peer over NVLink), the ranks rendezvous once,
and from then on the fused Lagrangian kernel itself pushes its K sums into every peer's mailbox and pulls the peers' --
no NCCL launch per step.  PyTorch provides the allocation and the rendezvous (plumbing); the exchange is in libikl.so.

    ex = PeerExchange()                       # after init_process_group("nccl")
    viol, loss, dtheta = fused.lagrangian_forward_backward(spec, theta, q_des, obst, lam, exchange=ex)   # viol: GLOBAL sums

    ar = PeerAllReduce(max_elems)             # the same idea for a flat fp32 vector: the replicas' parameter gradients
    ar.all_reduce_(flat)                      # in-place sum over ranks, one kernel, no NCCL launch, CUDA-graph capturable
"""
import ctypes

import torch
import torch.distributed as dist

from . import _cabi


class PeerExchange(object):
  def __init__(self, group=None, device=None):
    if not (dist.is_available() and dist.is_initialized()):
      raise RuntimeError("PeerExchange needs an initialised torch.distributed process group")
    import torch.distributed._symmetric_memory as symm_mem
    self.group = group if group is not None else dist.group.WORLD
    self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
    if self.world > _cabi.MAX_RANKS:
      raise ValueError("the in-kernel exchange serves up to %d ranks of one NVSwitch box, got %d" % (_cabi.MAX_RANKS, self.world))
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    nbytes = int(_cabi.load().ikl_exchange_bytes())
    self.buf = symm_mem.empty((nbytes + 7) // 8, dtype=torch.int64, device=dev)
    self.buf.zero_()
    self.handle = symm_mem.rendezvous(self.buf, self.group)
    ptrs = [int(p) for p in self.handle.buffer_ptrs]
    if len(ptrs) != self.world or ptrs[self.rank] != self.buf.data_ptr():
      raise RuntimeError("symmetric-memory rendezvous returned unexpected peer pointers")
    torch.cuda.synchronize(dev)
    dist.barrier(self.group)                    # every mailbox is zeroed before anybody's first launch can write into it
    self.c = _cabi.IklExchange()
    self.c.rank, self.c.world = self.rank, self.world
    for r in range(self.world):
      self.c.peers[r] = ptrs[r]

  def byref(self):
    return ctypes.byref(self.c)


class PeerAllReduce(object):
  """In-place sum over ranks of a flat fp32 CUDA vector of up to `max_elems` elements as ONE kernel over NVLink peer memory
  (include/ikl.h: ikl_allreduce_sum_f32).  Used for the parameter gradients of the data-parallel train step: no NCCL launch,
  capturable in the step's CUDA graph, bit-identical result on every rank."""

  def __init__(self, max_elems, group=None, device=None):
    if not (dist.is_available() and dist.is_initialized()):
      raise RuntimeError("PeerAllReduce needs an initialised torch.distributed process group")
    import torch.distributed._symmetric_memory as symm_mem
    self.group = group if group is not None else dist.group.WORLD
    self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
    if self.world > _cabi.MAX_RANKS:
      raise ValueError("the peer all-reduce serves up to %d ranks of one NVSwitch box, got %d" % (_cabi.MAX_RANKS, self.world))
    self.max_elems = int(max_elems)
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    self.device = dev
    nbytes = int(_cabi.load().ikl_allreduce_bytes(self.max_elems, self.world))
    if nbytes <= 0:
      raise ValueError("bad max_elems / world: %r / %r" % (max_elems, self.world))
    self.buf = symm_mem.empty((nbytes + 7) // 8, dtype=torch.int64, device=dev)
    self.buf.zero_()
    self.handle = symm_mem.rendezvous(self.buf, self.group)
    ptrs = [int(p) for p in self.handle.buffer_ptrs]
    if len(ptrs) != self.world or ptrs[self.rank] != self.buf.data_ptr():
      raise RuntimeError("symmetric-memory rendezvous returned unexpected peer pointers")
    torch.cuda.synchronize(dev)
    dist.barrier(self.group)                    # every mailbox is zeroed before anybody's first launch can write into it
    self.c = _cabi.IklExchange()
    self.c.rank, self.c.world = self.rank, self.world
    for r in range(self.world):
      self.c.peers[r] = ptrs[r]

  def all_reduce_(self, flat):
    """flat: contiguous 1-D float32 CUDA tensor with the SAME numel on every rank (<= max_elems); summed in place."""
    if not (isinstance(flat, torch.Tensor) and flat.is_cuda and flat.dtype == torch.float32 and flat.is_contiguous()):
      raise TypeError("all_reduce_ takes a contiguous float32 CUDA tensor")
    n = flat.numel()
    if n > self.max_elems:
      raise ValueError("%d elements exceed this mailbox's capacity of %d" % (n, self.max_elems))
    with torch.cuda.device(flat.device):
      st = _cabi.load().ikl_allreduce_sum_f32(ctypes.c_void_p(flat.data_ptr()), n, ctypes.byref(self.c), self.max_elems,
                                             ctypes.c_void_p(torch.cuda.current_stream(flat.device).cuda_stream))
    _cabi.check(st, "ikl_allreduce_sum_f32")
    return flat
