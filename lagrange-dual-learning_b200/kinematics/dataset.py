"""Drop-in for the reference's kinematics/dataset.py: IkDataset with the same constructor, the same item dictionary
(input_tensor (20,), joint_theta (J,), q_ee_desired (3,), optional dynamic_obstacles (4,4)) and the same cache file
format (kinematics/dataset.py:32-50,143-155).

Generation follows the reference's procedure and consumes torch's global CPU generator in the same order
(dataset.py:54-121): uniform joint angles inside the limits, forward kinematics for the targets, then rejection
sampling -- of obstacle centres for dynamic configurations, of joint angles for static ones -- with the reference's
collision rule (a joint tip or the arm base within `radius` of the centre, `<=`).  Differences, none of which change
what is sampled: forward kinematics runs in the CUDA kernel (ikl_fk_forward) on whole blocks of candidates instead of
once per retry; the random stream is drawn in blocks (torch's CPU uniform_ is a serial fill, so a block equals the
reference's sequence of small draws; the generator is re-wound to exactly the position the reference would leave it in);
and the per-candidate collision tests are plain float32 arithmetic instead of 0-d tensor ops.
"""
import math
import os

import numpy as np
import torch
from torch.utils.data import Dataset

from kinematics.forward_kinematics import ForwardKinematicsThreeLinkTorch, ForwardKinematicsEightLinkTorch
from kinematics.penalties import no_joint_collisions_circle, no_joint_collisions_rectangle  # noqa: F401  (API parity)

_F = np.float32


class _UniformStream(object):
  """Rows of `width` uniforms in [lo, hi) from torch's global CPU generator: drawn in blocks, handed out in order."""

  def __init__(self, lo, hi, width, block=4096):
    self.lo, self.hi, self.width, self.block = lo, hi, width, block
    self.state0 = torch.get_rng_state()
    self.buf = np.zeros((0, width), dtype=_F)
    self.pos = 0
    self.taken = 0

  def next(self):
    if self.pos == len(self.buf):
      self.buf = torch.empty(self.block * self.width).uniform_(self.lo, self.hi).numpy().reshape(self.block, self.width)
      self.pos = 0
    row = self.buf[self.pos]
    self.pos += 1
    self.taken += 1
    return row

  def give_back(self, n):
    """The caller did not use the last n rows it was handed."""
    self.taken -= n

  def close(self):
    """Leave the global generator exactly where a draw-by-draw consumer of `taken` rows would have left it."""
    torch.set_rng_state(self.state0)
    if self.taken:
      torch.empty(self.taken * self.width).uniform_(self.lo, self.hi)
    return self.taken


def _hits_circle_f32(tips_xy, cx, cy, radius32):
  """Reference rule (penalties.py:83-86) in float32: any tip with sqrt((qx-cx)^2 + (qy-cy)^2) <= radius."""
  dx = tips_xy[:, 0] - cx
  dy = tips_xy[:, 1] - cy
  return bool(np.any(np.sqrt(dx * dx + dy * dy) <= radius32))


class IkDataset(Dataset):
  """A torch `Dataset` for inverse kinematics tasks."""

  def __init__(self, N, J, config, seed=0, cache_save_path=None):
    """N examples for a J-link arm under the task `config` (see ikl.configs / resources of the reference).

    If `cache_save_path` exists the dataset is loaded from it, otherwise it is generated (and written there when a
    path is given) -- the file holds the four tensors random_theta, random_ee, random_obstacles, tensor_template.
    """
    super(IkDataset, self).__init__()
    self.N = N
    self.J = J
    self.config = config
    if cache_save_path is not None and os.path.exists(cache_save_path):
      print("[IkDataset] Found cached dataset at {}, loading from there!".format(cache_save_path))
      self._from_dict(torch.load(cache_save_path, map_location="cpu"))
    else:
      print("[IkDataset] Initializing dataset from scratch...")
      self.initialize(N, J, config, seed=seed)
      if cache_save_path is not None:
        print("[IkDataset] Saving the dataset to {}".format(cache_save_path))
        torch.save(self.as_dict(), cache_save_path)

  # -- (de)serialisation in the reference's cache format ---------------------------------------------------------
  def as_dict(self):
    return {"random_theta": self.random_theta, "random_ee": self.random_ee,
            "random_obstacles": self.random_obstacles, "tensor_template": self.tensor_template}

  def _from_dict(self, d):
    self.random_theta = d["random_theta"]
    self.random_ee = d["random_ee"]
    self.random_obstacles = d["random_obstacles"]
    self.tensor_template = d["tensor_template"]

  @classmethod
  def from_tensors(cls, config, tensors, J=3):
    """Wrap already generated tensors (e.g. a fixture or an on-device generator's output)."""
    self = cls.__new__(cls)
    Dataset.__init__(self)
    self.J, self.config = J, config
    self._from_dict(tensors)
    self.N = self.random_theta.shape[0]
    return self

  @classmethod
  def generate_on_device(cls, N, J, config, seed=0, device=None):
    """Benchmark-scale generation in one kernel launch (ikl_dataset_generate): same procedure and distribution as
    initialize(), Philox instead of torch's CPU stream, tensors stay on the GPU."""
    from ikl.device_data import generate_dataset
    assert J == 3
    tensors = generate_dataset(config, N, seed=seed, device=device, num_links=J)
    if int(tensors["exhausted"].item()) != 0:
      raise RuntimeError("dataset generation gave up on %d examples (an obstacle covers the arm base?)" % int(tensors["exhausted"].item()))
    return cls.from_tensors(config, tensors, J=J)

  # -- generation --------------------------------------------------------------------------------------------------
  def initialize(self, N, J, config, seed=0):
    torch.manual_seed(seed)
    lo = config["static_constraints"]["joint_limits"][0]
    hi = config["static_constraints"]["joint_limits"][1]
    print("[IkDataset] Joint angle limits:", lo, hi)
    self.random_theta = torch.empty(N, J).uniform_(lo, hi)
    fk = {3: ForwardKinematicsThreeLinkTorch, 8: ForwardKinematicsEightLinkTorch}[J]
    tips = fk(self.random_theta)                         # (x_J, ..., x_1), each (N, 3); CUDA kernel, CPU tensors back
    self.random_ee = tips[0]
    self.random_obstacles = None

    if config["dynamic_constraints"]["random_obstacles"]:
      self._place_dynamic_obstacles(tips, config)
    else:
      self._filter_static(fk, tips, config, lo, hi)

    # 20-float network input template: [ee_x, ee_y, limit_lo, limit_hi, 4 x (x, y, radius, -1 | zeros)]
    template = torch.zeros(20)
    template[2:4] = torch.Tensor(config["static_constraints"]["joint_limits"])
    for i, params in enumerate(config["static_constraints"]["obstacles"]):
      if params["type"] != "circle":
        raise NotImplementedError()
      template[4 + 4 * i:8 + 4 * i] = torch.Tensor([params["x"], params["y"], params["radius"], -1])
    self.tensor_template = template

  def _place_dynamic_obstacles(self, tips, config):
    n_obs = config["dynamic_constraints"]["random_obstacles_num"]
    template = config["dynamic_constraints"]["random_obstacles_template"]
    print("[IkDataset] Generating {} random obstacles for each example".format(n_obs))
    print("[IkDataset] Obstacle template:\n", template)
    N = len(self.random_ee)
    self.random_obstacles = torch.zeros(N, 4, 4)
    if template["type"] == "circle":
      self.random_obstacles[:, :, 2] = template["radius"]
    elif template["type"] == "rectangle":
      self.random_obstacles[:, :, 2] = template["width"]
      self.random_obstacles[:, :, 3] = template["height"]
      raise NotImplementedError("rectangle obstacles have no penalty in the reference either (penalties.py:27)")
    else:
      raise NotImplementedError("Unknown obstacle type {}".format(template["type"]))
    radius = template["radius"]
    radius32 = _F(radius)
    xy = np.stack([t[:, :2].numpy() for t in tips], axis=1).astype(_F)      # (N, J, 2), end effector first
    centres = np.zeros((N, 4, 2), dtype=_F)
    stream = _UniformStream(-1.2, 1.2, 2)
    for i in range(N):
      for o in range(n_obs):
        while True:
          cx, cy = stream.next()
          # arm base: the reference takes sqrt in double of the float32 sum and compares with the double radius
          base_hit = math.sqrt(float(cx * cx + cy * cy)) <= radius
          if not _hits_circle_f32(xy[i], cx, cy, radius32) and not base_hit:
            break
        centres[i, o] = (cx, cy)
    stream.close()
    self.random_obstacles[:, :, :2] = torch.from_numpy(centres)

  def _filter_static(self, fk, tips, config, lo, hi):
    obstacles = config["static_constraints"]["obstacles"]
    print("[IkDataset] Filtering dataset with {} static obstacles".format(len(obstacles)))
    if not obstacles:
      return
    N, J = self.random_theta.shape
    theta = self.random_theta.numpy()
    pose = np.stack([t.numpy() for t in tips], axis=1).astype(_F)           # (N, J, 3)
    stream = _UniformStream(lo, hi, J)
    # candidate replacement angles and their forward kinematics, block by block (one kernel launch per block)
    cand_theta, cand_pose, cursor = None, None, 0

    def next_candidate():
      nonlocal cand_theta, cand_pose, cursor
      if cand_theta is None or cursor == len(cand_theta):
        cand_theta = np.stack([stream.next().copy() for _ in range(1024)])
        cand_pose = np.stack([t.numpy() for t in fk(torch.from_numpy(cand_theta))], axis=1).astype(_F)
        cursor = 0
      cursor += 1
      return cand_theta[cursor - 1], cand_pose[cursor - 1]

    for i in range(N):
      for ob in obstacles:
        # NOTE: like the reference, only the CURRENT obstacle is re-checked after a re-draw (dataset.py:111-119)
        x32, y32, r32 = _F(ob["x"]), _F(ob["y"]), _F(ob["radius"])
        base_hit = math.sqrt(ob["x"] ** 2 + ob["y"] ** 2) <= ob["radius"]
        while _hits_circle_f32(pose[i, :, :2], x32, y32, r32) or base_hit:
          theta[i], pose[i] = next_candidate()
    # candidates were pre-drawn in blocks of 1024: hand the unused ones back before re-winding the generator
    stream.give_back(0 if cand_theta is None else len(cand_theta) - cursor)
    stream.close()
    self.random_theta = torch.from_numpy(theta)
    self.random_ee = torch.from_numpy(pose[:, 0, :].copy())

  # -- torch Dataset interface -----------------------------------------------------------------------------------
  def __len__(self):
    return self.random_theta.shape[0]

  def __getitem__(self, index):
    item = {}
    item["input_tensor"] = self.tensor_template.clone()
    item["input_tensor"][0:2] = self.random_ee[index, :2]
    item["joint_theta"] = self.random_theta[index]
    item["q_ee_desired"] = self.random_ee[index]
    if self.random_obstacles is not None:
      item["dynamic_obstacles"] = self.random_obstacles[index]
      item["input_tensor"][4:] = self.random_obstacles[index].flatten()
    return item
