"""Drop-in for the reference's kinematics/forward_kinematics.py: same names, same signatures, same
return conventions, computed by the CUDA kernels in libikl.so (ikl_fk_forward / ikl_fk_backward).

Reference interface replaced: kinematics/forward_kinematics.py:47-83 (ForwardKinematicsThreeLinkTorch),
:7-44 (the NumPy scalar helpers) and :86-110 (the eight-link function, which cannot run in the reference:
`batch_size` undefined at :99, `range()` without arguments at :104 -- here it is the function that loop spells out).
"""
import numpy as np
import torch

from ikl import ops as _ops
from utils.constants import Constants


def ForwardKinematicsTwoLink(theta):
  """EE pose (x, y, angle) of the 2-link arm for a length-2 array; object dtype so symbolic inputs work."""
  a0 = theta[0]
  a01 = a0 + theta[1]
  pose = np.empty(3, dtype=object)
  pose[0] = Constants.R2_LENGTH1 * np.cos(a0) + Constants.R2_LENGTH2 * np.cos(a01)
  pose[1] = Constants.R2_LENGTH1 * np.sin(a0) + Constants.R2_LENGTH2 * np.sin(a01)
  pose[2] = a01
  return pose


def ForwardKinematicsThreeLink(theta):
  """EE pose (x, y, angle) of the 3-link arm for a length-3 array; object dtype so symbolic inputs work."""
  lengths = (Constants.R3_LENGTH1, Constants.R3_LENGTH2, Constants.R3_LENGTH3)
  angle = 0
  pose = np.empty(3, dtype=object)
  pose[0] = pose[1] = 0
  for k in range(3):
    angle = angle + theta[k]
    pose[0] = pose[0] + lengths[k] * np.cos(angle)
    pose[1] = pose[1] + lengths[k] * np.sin(angle)
  pose[2] = angle
  return pose


def ForwardKinematicsThreeLinkTorch(theta):
  """theta (b, 3) -> (x3_ee, x2, x1), each (b, 3) = (x, y, cumulative angle); end effector FIRST.

  Differentiable w.r.t. theta.  Runs on the GPU (CPU inputs are moved there and back).
  """
  assert(len(theta.shape) == 2)
  assert(theta.shape[1] == 3)
  return _ops.fk_forward_kinematics(
      theta, (Constants.R3_LENGTH1, Constants.R3_LENGTH2, Constants.R3_LENGTH3))


def ForwardKinematicsTwoLinkTorch(theta):
  """theta (b, 2) -> (x2_ee, x1).  Not in the reference (its 2-link FK is NumPy-only); same conventions."""
  assert(len(theta.shape) == 2)
  assert(theta.shape[1] == 2)
  return _ops.fk_forward_kinematics(theta, (Constants.R2_LENGTH1, Constants.R2_LENGTH2))


def ForwardKinematicsEightLinkTorch(theta):
  """theta (b, 8) -> the eight tips, each (b, 3) = (x, y, cumulative angle), end effector FIRST.

  The reference's version cannot run (kinematics/forward_kinematics.py:99 uses an undefined `batch_size`, :104 calls
  `range()` without arguments); this is the function its loop body spells out -- link i adds R8_LENGTH at the cumulative
  angle, `links.reverse()` puts the end effector first (:105-110) -- computed by the same kernel as the 3-link arm.
  No reference output exists to pin it against: its parity is pinned only by the NumPy restatement in the tests.
  """
  assert(len(theta.shape) == 2)
  assert(theta.shape[1] == 8)
  return _ops.fk_forward_kinematics(theta, (Constants.R8_LENGTH,) * 8)
