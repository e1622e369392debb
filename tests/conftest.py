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
