"""The function-level CUDA drop-ins against the LIVE reference functions, side by side on the GPU box.

baseline/_ref/ (the byte-for-byte copy of the reference, tools/install_ref.py) travels with the snapshot, so the reference's
own kinematics/forward_kinematics.py and kinematics/penalties.py run here on the CPU (float32 and float64) while this repo's
modules of the same names run on the B200 through libikl.so -- on seeded random inputs beyond the committed golden vectors:
random shapes, broadcast obstacles, strided column views, slopes in either order / equal / zero, asymmetric and reversed
limits, exact kinks.  Bars (the ones of tests/test_gpu_parity.py):
  * joint-limit penalty: values and gradients BIT-IDENTICAL to the reference's float32 results, ties included;
  * circle penalty: values within 1 ulp(d) -- carried through t = d - r and the product -- of the reference's float32 result
    (torch's vectorised CPU sqrt is not correctly rounded everywhere; the kernel's is), at most 2 % of them different at all,
    and <= 1e-5 of slope * max(d, r) from its float64 result; gradients <= 2e-6 of scale against
    float64 away from |t| < 1e-6, exact ties carry torch's 1/2-split;
  * FK: angle column bit-identical, positions within 2e-6 (three fp32 sincos of ~1.5e-7 each) of the reference's float32 tips.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

SLOPES = [0.0, 0.005, 0.1, 0.5, 1.0, 2.5]


def _rand(gen, n, lo, hi):
  return (torch.rand(n, generator=gen, dtype=torch.float64) * (hi - lo) + lo).float()


def _cases(count, seed):
  rng = np.random.RandomState(seed)
  for i in range(count):
    yield i, torch.Generator().manual_seed(int(rng.randint(1 << 30))), rng


def test_joint_limit_dropin_equals_the_live_reference_bit_for_bit(cuda_device, live_reference):
  import kinematics.penalties as mine
  _, ref = live_reference
  launches = 0
  for i, gen, rng in _cases(40, 11):
    n = int(rng.choice([1, 2, 3, 31, 32, 33, 255, 1000, 4097]))
    inside, outside = float(rng.choice(SLOPES)), float(rng.choice(SLOPES))
    lo = float(rng.uniform(-3.2, 0.5))
    hi = lo + float(rng.uniform(-1.0, 5.0))                 # may be reversed: the formula uses |hi - lo|
    theta = _rand(gen, n * 3, -4.0, 4.0).reshape(n, 3)
    lim = torch.tensor([[lo, hi]], dtype=torch.float32)
    j = int(rng.randint(3))
    if n >= 3:                                              # on either limit and at the middle (sign(0) = 0)
      theta[0, j], theta[1, j], theta[2, j] = lim[0, 0], lim[0, 1], (lim[0, 0] + lim[0, 1]) / 2.0
    t_ref = theta.clone().requires_grad_(True)
    lim_b = lim.expand(n, -1)                               # trainer.py:170: a (1,2) tensor expanded over the batch
    want = ref.piecewise_joint_limit_penalty(t_ref[:, j], lim_b[:, 0], lim_b[:, 1], inside, outside)
    want.sum().backward()
    t_got = theta.to(cuda_device).requires_grad_(True)
    lim_g = lim.to(cuda_device).expand(n, -1)
    got = mine.piecewise_joint_limit_penalty(t_got[:, j], lim_g[:, 0], lim_g[:, 1], inside, outside)
    got.sum().backward()
    launches += 1
    assert got.is_cuda and got.shape == want.shape
    assert np.array_equal(got.detach().cpu().numpy(), want.detach().numpy()), "case %d: values" % i
    assert np.array_equal(t_got.grad.cpu().numpy(), t_ref.grad.numpy()), "case %d: gradient" % i
  assert launches == 40


def test_circle_dropin_against_the_live_reference(cuda_device, live_reference):
  import kinematics.penalties as mine
  _, ref = live_reference
  differing = total = 0
  for i, gen, rng in _cases(40, 23):
    n = int(rng.choice([1, 2, 5, 31, 64, 257, 1000, 4097]))
    inside, outside = float(rng.choice(SLOPES)), float(rng.choice(SLOPES))
    per_sample = bool(rng.randint(2))
    q = _rand(gen, n * 3, -3.0, 3.0).reshape(n, 3)          # (B,3) tips: the penalty sees column views
    ob = torch.zeros(n if per_sample else 1, 4, 4)
    ob[:, :, 0:2] = _rand(gen, ob.shape[0] * 8, -1.2, 1.2).reshape(-1, 4, 2)
    ob[:, :, 2] = _rand(gen, ob.shape[0] * 4, 0.05, 0.8).reshape(-1, 4)
    o = int(rng.randint(4))
    # exact tie: tip 0 ON the circle (a 3-4-5 triangle scaled by 1/4 is exact in fp32 around small centres)
    ob[0, o, 0], ob[0, o, 1], ob[0, o, 2] = 0.5, -0.25, 1.25
    q[0, 0], q[0, 1] = 0.5 + 0.75, -0.25 + 1.0
    outs = {}
    for tag, fn, dev, dt in (("f32", ref.piecewise_circle_penalty, "cpu", torch.float32),
                             ("f64", ref.piecewise_circle_penalty, "cpu", torch.float64),
                             ("cuda", mine.piecewise_circle_penalty, cuda_device, torch.float32)):
      qq = q.detach().clone().to(dev, dt).requires_grad_(True)
      oo = ob.to(dev, dt)
      p = fn(qq[:, 0], qq[:, 1], oo[:, o, 0], oo[:, o, 1], oo[:, o, 2], inside, outside)      # trainer.py:233-238
      p.sum().backward()
      outs[tag] = (p.detach().cpu().double().numpy(), qq.grad.cpu().double().numpy())
    (v32, g32), (v64, g64), (vc, gc) = outs["f32"], outs["f64"], outs["cuda"]
    assert vc.shape == v32.shape == (n,)
    d64 = np.sqrt((q[:, 0].double().numpy() - ob[:, o, 0].double().numpy()) ** 2 + (q[:, 1].double().numpy() - ob[:, o, 1].double().numpy()) ** 2)
    top = max(inside, outside)
    r64 = ob[:, o, 2].double().numpy()
    t64 = d64 - r64
    # one ulp of d, carried through t = d - r (which rounds again) and the product (which rounds again)
    ulp = lambda x: np.spacing(np.abs(x).astype(np.float32)).astype(np.float64)
    assert np.all(np.abs(vc - v32) <= 1.01 * top * (ulp(d64) + ulp(t64)) + ulp(v32)), \
        "case %d: value beyond 1 ulp(d) of the reference's float32 result" % i
    differing += int((vc != v32).sum())
    total += n
    scale = top * max(float(d64.max()), float(r64.max()))           # the magnitudes that enter t = d - r
    assert np.all(np.abs(vc - v64) <= 1e-5 * scale), "case %d: value vs float64" % i
    away = np.abs(t64) > 1e-6
    assert np.all(np.abs(gc - g64)[away] <= 2e-6 * max(top, 1e-30)), "case %d: gradient vs float64" % i
    assert float(np.abs(gc[:, 2]).sum()) == 0.0                     # the angle column takes no gradient
    # the exact tie: the averaged slope, like torch's MaximumBackward (sample 0; t == 0 in float32 and float64 alike)
    assert t64[0] == 0.0
    np.testing.assert_allclose(gc[0], g32[0], rtol=1e-6, atol=1e-7)
  assert differing <= 0.02 * total, "%d of %d values differ from torch's CPU float32 result" % (differing, total)


def test_fk_dropin_against_the_live_reference(cuda_device, live_reference):
  import kinematics.forward_kinematics as mine
  ref, _ = live_reference
  for i, gen, rng in _cases(24, 37):
    n = int(rng.choice([1, 2, 7, 32, 100, 1023, 4097]))
    scale = float(rng.choice([0.5, 2.5, 40.0, 3.0e4]))
    theta = _rand(gen, n * 3, -scale, scale).reshape(n, 3)
    w = _rand(gen, 9, -1.0, 1.0).reshape(3, 3)
    a = theta.clone().requires_grad_(True)
    want = ref.ForwardKinematicsThreeLinkTorch(a)                    # (x3_ee, x2, x1), forward_kinematics.py:83
    sum((x * w[k]).sum() for k, x in enumerate(want)).backward()
    b = theta.to(cuda_device).requires_grad_(True)
    got = mine.ForwardKinematicsThreeLinkTorch(b)
    wg = w.to(cuda_device)
    sum((x * wg[k]).sum() for k, x in enumerate(got)).backward()
    assert isinstance(got, tuple) and len(got) == 3
    for k in range(3):
      g, r = got[k].detach().cpu().numpy(), want[k].detach().numpy()
      assert g.shape == r.shape == (n, 3)
      assert np.array_equal(g[:, 2], r[:, 2]), "case %d tip %d: cumulative angle" % (i, k)
      assert np.abs(g[:, :2] - r[:, :2]).max() <= 2e-6, "case %d tip %d: position" % (i, k)
    # d/dtheta: sums of +-sin / cos of the same float32 angles weighted by |w| <= 1 -- at most 6 terms of ~1.5e-7 each, plus
    # the round-off of the reference's own float32 accumulation
    assert np.abs(b.grad.cpu().numpy() - a.grad.numpy()).max() <= 5e-6, "case %d: gradient" % i


def test_the_references_own_test_file_passes_over_the_dropins(cuda_device):
  """test/test_training_utils.py of the reference, UNMODIFIED and executed by file path from baseline/_ref/: its
  `from kinematics.penalties import ...` (line 4) resolves to this repo's drop-in, so its two known-answer tests run on
  libikl.so (CPU tensors in and out, as the reference's tests pass them; the launch counter proves the kernels ran)."""
  import importlib.util
  import inspect
  import os
  from conftest import PKG, reference_root
  from ikl import _cabi
  root = reference_root()
  if root is None:
    pytest.skip("no copy of the reference (baseline/_ref or /root/reference)")
  path = os.path.join(root, "test", "test_training_utils.py")
  spec = importlib.util.spec_from_file_location("_ref_test_training_utils", path)
  mod = importlib.util.module_from_spec(spec)
  spec.loader.exec_module(mod)
  src = os.path.realpath(inspect.getsourcefile(mod.piecewise_circle_penalty))
  assert src.startswith(os.path.realpath(PKG) + os.sep), "the reference's test must be exercising the drop-in, not %s" % src
  tests = [(n, f) for n, f in sorted(vars(mod).items()) if n.startswith("test_") and callable(f)]
  assert [n for n, _ in tests] == ["test_piecewise_circle_penalty", "test_piecewise_joint_limit_penalty"]
  before = _cabi.launch_count()
  for _, fn in tests:
    fn()
  assert _cabi.launch_count() - before >= 2, "at least one libikl.so launch per penalty call"
