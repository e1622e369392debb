"""Device-side callers of the hot path: dataset generation at benchmark scale and evaluation statistics.

  generate_dataset   reference kinematics/dataset.py:52-121 (distributional equivalent; one Philox stream per example)
  eval_stats         reference scripts/evaluation/visualize_ik_solutions.py:97-119
  input_tensor       the 20-float network input of kinematics/dataset.py:124-153, assembled for a whole batch
"""
import ctypes

import torch

from . import _cabi
from .fused import _batch_desc
from .spec import ConstraintSpec


def _dataset_spec(config, num_links=3):
  """Generation needs the obstacle geometry even when cfg["enforce_obstacles"] is false (the dataset code ignores that flag)."""
  dyn = config["dynamic_constraints"]["random_obstacles"] == True  # noqa: E712
  static = [(o["x"], o["y"], o["radius"]) for o in config["static_constraints"]["obstacles"]]
  n = int(config["dynamic_constraints"]["random_obstacles_num"]) if dyn else len(static)
  return ConstraintSpec(num_links=num_links, enforce_joint_limits=True, joint_limits=config["static_constraints"]["joint_limits"],
                        num_obstacles=n, dynamic=dyn, static_obstacles=() if dyn else static)


def generate_dataset(config, n, seed=0, device=None, num_links=3):
  """Returns a dict with the reference's cache keys: random_theta (n,3), random_ee (n,3), random_obstacles (n,4,4) or
  None, tensor_template (20,) -- all on `device`."""
  if not torch.cuda.is_available():
    raise RuntimeError("generate_dataset runs on the GPU (kinematics.dataset.IkDataset is the host-stream-compatible generator)")
  dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
  spec = _dataset_spec(config, num_links)
  lib = _cabi.load()
  theta = torch.empty(n, 3, dtype=torch.float32, device=dev)
  ee = torch.empty(n, 3, dtype=torch.float32, device=dev)
  obst = torch.empty(n, 4, 4, dtype=torch.float32, device=dev) if spec.dynamic else None
  exhausted = torch.zeros(1, dtype=torch.int32, device=dev)
  radius = float(config["dynamic_constraints"]["random_obstacles_template"]["radius"]) if spec.dynamic else 0.0
  with torch.cuda.device(dev):
    st = lib.ikl_dataset_generate(ctypes.byref(spec.c_struct), n, int(seed) & 0xFFFFFFFFFFFFFFFF, radius, ctypes.c_void_p(theta.data_ptr()),
                                  ctypes.c_void_p(ee.data_ptr()), ctypes.c_void_p(obst.data_ptr()) if obst is not None else None,
                                  ctypes.c_void_p(exhausted.data_ptr()), ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
  _cabi.check(st, "ikl_dataset_generate")
  template = torch.zeros(20)
  template[2:4] = torch.Tensor(config["static_constraints"]["joint_limits"])
  for i, ob in enumerate(config["static_constraints"]["obstacles"]):
    template[4 + 4 * i:8 + 4 * i] = torch.Tensor([ob["x"], ob["y"], ob["radius"], -1])
  return {"random_theta": theta, "random_ee": ee, "random_obstacles": obst, "tensor_template": template.to(dev), "exhausted": exhausted}


def input_tensor(tensors, start=0, stop=None):
  """(n,20) network inputs for examples [start, stop) of a generated dataset, on its device."""
  ee = tensors["random_ee"][start:stop]
  x = tensors["tensor_template"].to(ee.device).unsqueeze(0).repeat(ee.shape[0], 1)
  x[:, 0:2] = ee[:, :2]
  if tensors["random_obstacles"] is not None:
    x[:, 4:] = tensors["random_obstacles"][start:stop].reshape(ee.shape[0], 16)
  return x


def eval_stats(spec, theta, q_ee_desired, obstacles=None, threshold=1e-4, counts=None, pos_err_sum=None):
  """Accumulate the evaluation statistics of one batch.  Returns (counts (18,) int64 on device, pos_err_sum (1,) float64):
  counts[:K] per-constraint #(term > threshold), counts[16] #samples violating any joint limit, counts[17] any obstacle."""
  lib = _cabi.load()
  desc, keep = _batch_desc(spec, theta, q_ee_desired, obstacles)
  dev = theta.device
  if counts is None:
    counts = torch.zeros(18, dtype=torch.int64, device=dev)
  if pos_err_sum is None:
    pos_err_sum = torch.zeros(1, dtype=torch.float64, device=dev)
  with torch.cuda.device(dev):
    st = lib.ikl_eval_stats(ctypes.byref(spec.c_struct), ctypes.byref(desc), float(threshold), ctypes.c_void_p(counts.data_ptr()),
                            ctypes.c_void_p(pos_err_sum.data_ptr()), ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
  _cabi.check(st, "ikl_eval_stats")
  return counts, pos_err_sum
