"""Shared test plumbing: marker registration, sys.path for the drop-in tree, fixture loaders."""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "lagrange-dual-learning_b200")
FUSED_OVERLAY = os.path.join(PKG, "fused_overlay")     # the fused trainer mirror (`scripts` package); see INTEGRATION.md
GOLDEN = os.path.join(ROOT, "tests", "golden")
for p in (ROOT, FUSED_OVERLAY, PKG):
  if p not in sys.path:
    sys.path.insert(0, p)

CONFIG_NAMES = [
    "cfg_no_constraints",
    "cfg_joint_limits",
    "cfg_joint_limits_and_1obs_static",
    "cfg_joint_limits_and_4obs_static",
    "cfg_joint_limits_and_4obs_dynamic",
]


def pytest_configure(config):
  config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden_npz(name):
  return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def config_json(name):
  """One of the five standard task configs (built by ikl.configs; equality with the reference's
  resources/cfg_*.json is checked in test_oracle_pinned.py when /root/reference is mounted)."""
  from ikl.configs import standard_config
  return standard_config(name)


@pytest.fixture(scope="session")
def cuda_device():
  import torch
  if not torch.cuda.is_available():
    pytest.skip("no CUDA device")
  return torch.device("cuda:0")


def reference_root():
  """A byte-for-byte copy of the reference: baseline/_ref/ (tools/install_ref.py; travels to the GPU box) or, in the build
  container, /root/reference.  None when neither exists."""
  for root in (os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
    if os.path.exists(os.path.join(root, "kinematics", "penalties.py")):
      return root
  return None


def _load_reference_file(name, rel):
  import importlib.util
  spec = importlib.util.spec_from_file_location(name, os.path.join(reference_root(), rel))
  mod = importlib.util.module_from_spec(spec)
  spec.loader.exec_module(mod)
  return mod


@pytest.fixture(scope="session")
def live_reference():
  """(forward_kinematics module, penalties module) of the REFERENCE, executed by file path under private module names so
  that they cannot be confused with this repo's CUDA drop-ins of the same names.  forward_kinematics.py:4 imports
  utils.constants: the reference's own file is supplied under that name for the duration of the import and removed again."""
  import types
  root = reference_root()
  if root is None:
    pytest.skip("no copy of the reference (baseline/_ref or /root/reference)")
  consts = _load_reference_file("_ref_utils_constants", os.path.join("utils", "constants.py"))
  saved = {k: sys.modules.get(k) for k in ("utils", "utils.constants")}
  pkg = types.ModuleType("utils")
  pkg.constants = consts
  sys.modules["utils"], sys.modules["utils.constants"] = pkg, consts
  try:
    fk = _load_reference_file("_ref_forward_kinematics", os.path.join("kinematics", "forward_kinematics.py"))
  finally:
    for k, v in saved.items():
      if v is None:
        sys.modules.pop(k, None)
      else:
        sys.modules[k] = v
  pen = _load_reference_file("_ref_penalties", os.path.join("kinematics", "penalties.py"))
  for mod in (fk, pen):
    assert os.path.realpath(mod.__file__).startswith(os.path.realpath(root) + os.sep)
  return fk, pen
