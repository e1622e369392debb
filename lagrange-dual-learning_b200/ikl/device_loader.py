"""Batches straight from device-resident dataset tensors (no DataLoader workers, no per-example Python).

The reference iterates `DataLoader(IkDataset, batch_size, shuffle=True)` (scripts/trainer.py:155-156): every batch is
collated from `batch_size` calls of IkDataset.__getitem__ (kinematics/dataset.py:143-155), which is fine at its native
batch of 8 and hopeless at the batch sizes a B200 wants.  DeviceBatchLoader yields the same dictionaries (input_tensor,
joint_theta, q_ee_desired[, dynamic_obstacles]) by indexing the dataset's tensors with one on-device permutation per
epoch; under torch.distributed each rank takes its strided share of every global batch.
"""
import torch

from . import dist as ikl_dist


class DeviceBatchLoader(object):
  def __init__(self, dataset, batch_size, shuffle=True, device=None, rank=None, world=None, seed=0, planes=False):
    self.ds = dataset
    self.batch_size = int(batch_size)           # GLOBAL batch size
    self.shuffle = shuffle
    self.rank = ikl_dist.rank() if rank is None else rank
    self.world = ikl_dist.world_size() if world is None else world
    dev = dataset.random_theta.device if device is None else torch.device(device)
    self.device = dev
    self.theta = dataset.random_theta.to(dev)
    self.ee = dataset.random_ee.to(dev)
    self.obst = None if dataset.random_obstacles is None else dataset.random_obstacles.to(dev)
    self.template = dataset.tensor_template.to(dev)
    # planes=True: keep targets / obstacles as structure-of-arrays planes so that every batch comes out in the fused
    # kernels' native layout (the gather by index writes a fresh tensor anyway, so this costs nothing extra)
    self.planes = planes
    if planes:
      self.ee_planes = self.ee.t().contiguous()                                             # (3, N)
      self.obst_planes = None if self.obst is None else self.obst[:, :, :3].permute(1, 2, 0).contiguous()   # (4, 3, N)
    self.gen = torch.Generator(device=dev)
    self.gen.manual_seed(seed)                  # same seed on every rank: all ranks draw the same permutation
    self.epoch = 0

  def __len__(self):
    n = self.theta.shape[0]
    return (n + self.batch_size - 1) // self.batch_size

  def batch(self, idx):
    """The reference's item dictionary, for a whole index tensor at once."""
    if self.planes:
      return self._batch_planes(idx)
    ee = self.ee[idx]
    x = self.template.unsqueeze(0).repeat(idx.numel(), 1)
    x[:, 0:2] = ee[:, :2]
    out = {"joint_theta": self.theta[idx], "q_ee_desired": ee}
    if self.obst is not None:
      ob = self.obst[idx]
      out["dynamic_obstacles"] = ob
      x[:, 4:] = ob.reshape(idx.numel(), 16)
    out["input_tensor"] = x
    return out

  def _batch_planes(self, idx):
    n = idx.numel()
    ee = self.ee_planes[:, idx].t()                       # (n, 3) view of a (3, n) buffer: one plane per column
    x = self.template.unsqueeze(0).repeat(n, 1)
    x[:, 0:2] = ee[:, :2]
    out = {"joint_theta": self.theta[idx], "q_ee_desired": ee}
    if self.obst_planes is not None:
      ob = self.obst_planes[:, :, idx].permute(2, 0, 1)   # (n, 4, 3) view of a (4, 3, n) buffer
      out["dynamic_obstacles"] = ob
      x.view(n, 5, 4)[:, 1:, :3] = ob                     # slots 4..19 of the input: 4 x (x, y, radius, 0)
      x.view(n, 5, 4)[:, 1:, 3] = 0
    out["input_tensor"] = x
    return out

  def __iter__(self):
    n = self.theta.shape[0]
    order = torch.randperm(n, generator=self.gen, device=self.device) if self.shuffle else torch.arange(n, device=self.device)
    self.epoch += 1
    for s in range(0, n, self.batch_size):
      idx = order[s:s + self.batch_size][self.rank::self.world]      # this rank's share of the global batch
      yield self.batch(idx)
