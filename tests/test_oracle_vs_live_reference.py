"""Pin the oracle against the LIVE reference functions on random inputs (beyond the committed golden vectors).

The reference's own kinematics/forward_kinematics.py and kinematics/penalties.py are executed by FILE PATH -- from
baseline/_ref/ (the byte-for-byte copy tools/install_ref.py makes; it travels to the GPU box) or, in the build container,
from /root/reference -- under private module names, so they cannot be confused with this repo's CUDA drop-ins of the same
names.  hypothesis draws shapes, slopes (either order, equal, zero), limits (asymmetric, reversed) and inputs, including
exact kinks; the oracle must reproduce values AND gradients bit for bit in float32 and float64.  CPU only."""
import os
import sys

import numpy as np
import pytest
import torch
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from conftest import ROOT
from oracle import ik_oracle as O


@pytest.fixture(scope="module")
def ref(live_reference):
  return live_reference


# derandomize: the same examples on every run (a CI suite must not depend on the draw); no example database on disk
_SETTINGS = dict(max_examples=120, deadline=None, derandomize=True, database=None,
                 suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])
_slopes = st.sampled_from([0.0, 0.005, 0.1, 0.5, 1.0, 2.5])
_dtypes = st.sampled_from([torch.float32, torch.float64])


def _rand(seed, n, lo, hi, dtype):
  g = torch.Generator().manual_seed(seed)
  return (torch.rand(n, generator=g, dtype=torch.float64) * (hi - lo) + lo).to(dtype)


def _same_bits(a, b, what):
  a, b = a.detach().numpy(), b.detach().numpy()
  assert a.dtype == b.dtype and a.shape == b.shape, what
  assert np.array_equal(a, b, equal_nan=True), "%s: %d of %d elements differ" % (what, int((a != b).sum()), a.size)


@settings(**_SETTINGS)
@given(seed=st.integers(0, 2**31 - 1), n=st.integers(1, 257), dtype=_dtypes, scale=st.sampled_from([0.5, 2.5, 40.0, 3.0e4]))
def test_fk_three_link_bitwise(ref, seed, n, dtype, scale):
  fk, _ = ref
  theta = _rand(seed, n * 3, -scale, scale, dtype).reshape(n, 3)
  a = theta.clone().requires_grad_(True)
  b = theta.clone().requires_grad_(True)
  want = fk.ForwardKinematicsThreeLinkTorch(a)              # forward_kinematics.py:47-83, end effector first
  got = O.forward_kinematics(b)
  assert len(want) == len(got) == 3
  w = _rand(seed + 1, 9, -1.0, 1.0, dtype).reshape(3, 3)
  for i, (x, y) in enumerate(zip(want, got)):
    _same_bits(y, x, "tip %d" % i)
  sum((x * w[i]).sum() for i, x in enumerate(want)).backward()
  sum((y * w[i]).sum() for i, y in enumerate(got)).backward()
  # the reference accumulates through slice assignments, the oracle through stack: the same per-element sums in a
  # different order of accumulation -- equal to round-off, not bit for bit
  tol = 1e-5 if dtype == torch.float32 else 1e-13
  np.testing.assert_allclose(b.grad.numpy(), a.grad.numpy(), rtol=tol, atol=tol)


@settings(**_SETTINGS)
@given(seed=st.integers(0, 2**31 - 1), n=st.integers(1, 300), dtype=_dtypes, inside=_slopes, outside=_slopes,
       scalar_obstacle=st.booleans(), on_circle=st.booleans())
def test_circle_penalty_bitwise_with_gradients(ref, seed, n, dtype, inside, outside, scalar_obstacle, on_circle):
  _, pen = ref
  m = 1 if scalar_obstacle else n
  jx, jy = _rand(seed, n, -3.0, 3.0, dtype), _rand(seed + 1, n, -3.0, 3.0, dtype)
  ox, oy = _rand(seed + 2, m, -1.2, 1.2, dtype), _rand(seed + 3, m, -1.2, 1.2, dtype)
  radius = _rand(seed + 4, m, 0.05, 0.8, dtype)
  if on_circle:                                            # exact ties: the point ON the circle (3-4-5 triangles are exact)
    jx[0], jy[0] = ox[0] + 0.75, oy[0] + 1.0
    radius[0] = torch.sqrt((jx[0] - ox[0]) ** 2 + (jy[0] - oy[0]) ** 2)
  outs = []
  for fn in (pen.piecewise_circle_penalty, O.circle_penalty):
    x, y = jx.clone().requires_grad_(True), jy.clone().requires_grad_(True)
    p = fn(x, y, ox, oy, radius, inside, outside)           # penalties.py:5-17
    p.sum().backward()
    outs.append((p, x.grad, y.grad))
  for k, what in enumerate(("value", "d/djx", "d/djy")):
    _same_bits(outs[1][k], outs[0][k], what)


@settings(**_SETTINGS)
@given(seed=st.integers(0, 2**31 - 1), n=st.integers(1, 300), dtype=_dtypes, inside=_slopes, outside=_slopes,
       lo=st.floats(-3.2, 0.5), width=st.floats(-1.0, 5.0), kinks=st.booleans())
def test_joint_limit_penalty_bitwise_with_gradients(ref, seed, n, dtype, inside, outside, lo, width, kinks):
  _, pen = ref
  hi = lo + width                                          # width < 0: reversed limits, |hi - lo| is what the formula uses
  theta = _rand(seed, n, -4.0, 4.0, dtype)
  lim_lo, lim_hi = torch.tensor([lo], dtype=dtype), torch.tensor([hi], dtype=dtype)
  if kinks and n >= 3:                                     # on a limit, on the other one, at the middle
    theta[0], theta[1], theta[2] = lim_lo[0], lim_hi[0], (lim_lo[0] + lim_hi[0]) / 2.0
  outs = []
  for fn in (pen.piecewise_joint_limit_penalty, O.joint_limit_penalty):
    t = theta.clone().requires_grad_(True)
    p = fn(t, lim_lo.expand(n), lim_hi.expand(n), inside, outside)     # penalties.py:44-56; expanded limits as trainer.py:170,202
    p.sum().backward()
    outs.append((p, t.grad))
  _same_bits(outs[1][0], outs[0][0], "value")
  _same_bits(outs[1][1], outs[0][1], "d/dtheta")


def test_collision_predicates_follow_the_reference(ref):
  """no_joint_collisions_circle / _rectangle (penalties.py:58-92): collision iff a tip OR the origin is inside or ON the
  obstacle (`<=`).  The drop-in's predicates are host logic (kinematics/penalties.py of this repo) -- compared with the
  reference's on random cases that include exact touches."""
  _, pen = ref
  import importlib
  sys.path.insert(0, os.path.join(ROOT, "lagrange-dual-learning_b200"))
  try:
    mine = importlib.import_module("kinematics.penalties")
  finally:
    sys.path.pop(0)
  assert os.path.realpath(mine.__file__).startswith(os.path.realpath(os.path.join(ROOT, "lagrange-dual-learning_b200")))
  rng = np.random.RandomState(7)
  seen = {True: 0, False: 0}
  for _ in range(400):
    tips = [torch.tensor(rng.uniform(-3, 3, size=3), dtype=torch.float32) for _ in range(3)]
    ox, oy = [float(v) for v in rng.uniform(-1.5, 1.5, size=2)]
    r = float(rng.uniform(0.05, 1.0))
    if rng.rand() < 0.2:                                   # the origin exactly on the circle: 3-4-5
      ox, oy, r = 0.75, 1.0, 1.25
    elif rng.rand() < 0.2:                                 # a tip exactly on the circle
      ox, oy, r = 2.0, -1.0, 0.625
      tips[1][0], tips[1][1] = ox + 0.375, oy + 0.5
    want = pen.no_joint_collisions_circle(tips, ox, oy, {"radius": r})
    got = mine.no_joint_collisions_circle(tips, ox, oy, {"radius": r})
    assert bool(want) == bool(got), (tips, ox, oy, r)
    seen[bool(want)] += 1
    w, h = [float(v) for v in rng.uniform(0.1, 2.0, size=2)]
    if rng.rand() < 0.2:                                   # a tip exactly on the rectangle's edge
      tips[2][0], tips[2][1] = ox + w, oy
    box = {"width": w, "height": h}
    assert bool(pen.no_joint_collisions_rectangle(tips, ox, oy, box)) == bool(mine.no_joint_collisions_rectangle(tips, ox, oy, box))
  assert seen[True] > 50 and seen[False] > 50
  with pytest.raises(AssertionError):
    mine.no_joint_collisions_circle([], 0.0, 0.0, {})
  with pytest.raises(AssertionError):
    mine.no_joint_collisions_rectangle([], 0.0, 0.0, {"width": 1.0})
