"""Function-level CUDA ops behind the reference's free functions (level (i) of the drop-in boundary):

  fk_forward_kinematics   <- kinematics/forward_kinematics.py:47-83  ForwardKinematicsThreeLinkTorch (J = 2 or 3)
  circle_penalty          <- kinematics/penalties.py:5-17             piecewise_circle_penalty
  joint_limit_penalty     <- kinematics/penalties.py:44-56            piecewise_joint_limit_penalty

Each is a torch.autograd.Function over one C-ABI launch.  They accept what the reference's callers pass:
non-contiguous column views, expand()-ed limits and broadcasting (scripts/trainer.py:201-204, :235-238;
test/test_training_utils.py:7-41).  CPU tensors (kinematics/dataset.py:68 calls FK on CPU) are moved to the
GPU, computed there and moved back -- the arithmetic always runs in the CUDA kernels; without a CUDA device
the call raises.
"""
import ctypes

import torch

from . import _cabi


def _device_for(*tensors):
  for t in tensors:
    if isinstance(t, torch.Tensor) and t.is_cuda:
      return t.device
  if not torch.cuda.is_available():
    raise RuntimeError("ikl ops need a CUDA device: there is no CPU implementation of the IK kernels")
  return torch.device("cuda", torch.cuda.current_device())


def _require_f32(t, name):
  """The kernels compute in float32 only.  The reference's functions are dtype-generic (a float64 tensor is evaluated in
  float64 by torch); silently narrowing such an input would be different arithmetic, so it is refused instead."""
  if isinstance(t, torch.Tensor) and t.dtype != torch.float32:
    raise TypeError("%s is %s: the CUDA drop-ins compute in float32 only (include/ikl.h); cast explicitly with .float() "
                    "if float32 arithmetic is what you want" % (name, t.dtype))


def _stream(dev):
  return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


class _on_device(object):
  """`with torch.cuda.device(dev)` only when dev is not already current: an unchanged scripts/trainer.py makes ~30 of these
  calls per step and is host-bound, so the common case must not pay for a device switch that does nothing."""
  __slots__ = ("guard",)

  def __init__(self, dev):
    self.guard = None if dev.index is None or dev.index == torch.cuda.current_device() else torch.cuda.device(dev)

  def __enter__(self):
    if self.guard is not None:
      self.guard.__enter__()

  def __exit__(self, *exc):
    if self.guard is not None:
      return self.guard.__exit__(*exc)
    return False


def _vp(t):
  return None if t is None else ctypes.c_void_p(t.data_ptr())


def _lengths(link_lengths, J):
  vals = [1.0] * J if link_lengths is None else [float(v) for v in link_lengths]
  return (ctypes.c_float * J)(*vals)


# ---- forward kinematics --------------------------------------------------------------------------------------
class _ForwardKinematics(torch.autograd.Function):
  @staticmethod
  def forward(ctx, theta, link_lengths):
    lib = _cabi.load()
    B, J = theta.shape
    tips = torch.empty(J, B, 3, dtype=torch.float32, device=theta.device)
    with _on_device(theta.device):
      st = lib.ikl_fk_forward(_vp(theta), theta.stride(0), theta.stride(1), B, J, _lengths(link_lengths, J), _vp(tips),
                              _stream(theta.device))
    _cabi.check(st, "ikl_fk_forward")
    ctx.save_for_backward(theta)
    ctx.link_lengths = link_lengths
    return tuple(tips[k] for k in range(J))

  @staticmethod
  def backward(ctx, *grads):
    (theta,) = ctx.saved_tensors
    lib = _cabi.load()
    B, J = theta.shape
    g = torch.stack([torch.zeros(B, 3, dtype=torch.float32, device=theta.device) if gk is None else gk.to(torch.float32)
                     for gk in grads]).contiguous()
    dtheta = torch.empty(B, J, dtype=torch.float32, device=theta.device)
    with _on_device(theta.device):
      st = lib.ikl_fk_backward(_vp(theta), theta.stride(0), theta.stride(1), B, J, _lengths(ctx.link_lengths, J), _vp(g),
                               _vp(dtheta), _stream(theta.device))
    _cabi.check(st, "ikl_fk_backward")
    return dtheta, None


def fk_forward_kinematics(theta, link_lengths=None):
  """theta (B,J) -> tuple of J (B,3) tensors, END EFFECTOR FIRST, rows (x, y, cumulative angle)."""
  assert len(theta.shape) == 2                      # forward_kinematics.py:57
  _require_f32(theta, "theta")
  src_device = theta.device
  dev = _device_for(theta)
  th = theta.to(device=dev)
  tips = _ForwardKinematics.apply(th, None if link_lengths is None else tuple(link_lengths))
  if src_device != dev:
    tips = tuple(t.to(device=src_device) for t in tips)
  return tips


# ---- elementwise penalties -----------------------------------------------------------------------------------
def _flat_view(t, shape):
  """Broadcast t to `shape` and describe it as (tensor_to_keep_alive, data_ptr, element stride) over prod(shape) elements."""
  if t.dim() == 1 and len(shape) == 1 and t.shape[0] == shape[0]:
    return t, (t.stride(0) if shape[0] > 1 else 0)          # the trainer's case: a column view or an expand()-ed limit, as is
  e = t.expand(shape)
  n = e.numel()
  if n <= 1 or all(s == 0 for s, d in zip(e.stride(), e.shape) if d > 1):
    return e, 0
  if e.dim() == 1:
    return e, e.stride(0)
  if e.is_contiguous():
    return e, 1
  e = e.contiguous()
  return e, 1


def _prep(dev, *ts):
  # fast path (every call an unchanged scripts/trainer.py makes): float32 tensors already on the device, one shape
  t0 = ts[0]
  if isinstance(t0, torch.Tensor):
    shape = t0.shape
    for t in ts:
      if not (isinstance(t, torch.Tensor) and t.dtype == torch.float32 and t.device == dev and t.shape == shape):
        break
    else:
      return ts, shape
  out = []
  for i, t in enumerate(ts):
    if not isinstance(t, torch.Tensor):
      t = torch.tensor(t, dtype=torch.float32)
    _require_f32(t, "operand %d" % i)
    out.append(t.to(device=dev))
  shape = torch.broadcast_shapes(*[t.shape for t in out])
  return out, shape


def _reduce_to(g, shape):
  """Sum a broadcast-shaped gradient back to an operand's shape."""
  return g.sum_to_size(shape) if tuple(g.shape) != tuple(shape) else g


class _CirclePenalty(torch.autograd.Function):
  @staticmethod
  def forward(ctx, jx, jy, ox, oy, radius, inside_slope, outside_slope):
    lib = _cabi.load()
    shape = jx.shape if jx.shape == jy.shape == ox.shape == oy.shape == radius.shape else \
        torch.broadcast_shapes(jx.shape, jy.shape, ox.shape, oy.shape, radius.shape)
    views = [_flat_view(t, shape) for t in (jx, jy, ox, oy, radius)]
    out = torch.empty(shape, dtype=torch.float32, device=jx.device)
    args = []
    for v, s in views:
      args += [_vp(v), s]
    with _on_device(jx.device):
      st = lib.ikl_circle_penalty_forward(*args, out.numel(), inside_slope, outside_slope, _vp(out), _stream(jx.device))
    _cabi.check(st, "ikl_circle_penalty_forward")
    ctx.save_for_backward(jx, jy, ox, oy, radius)
    ctx.slopes = (inside_slope, outside_slope)
    return out

  @staticmethod
  def backward(ctx, grad_out):
    lib = _cabi.load()
    operands = ctx.saved_tensors
    shape = grad_out.shape
    views = [_flat_view(t, shape) for t in operands]
    # the gradient of the `.sum()` the trainer applies (scripts/trainer.py:239) is ONE value expand()-ed over the batch: it is
    # handed to the kernel with element stride 0 instead of being materialised
    g, s_g = _flat_view(grad_out if grad_out.dtype == torch.float32 else grad_out.to(torch.float32), shape)
    need = ctx.needs_input_grad[:5]
    outs = [torch.empty(shape, dtype=torch.float32, device=g.device) if n else None for n in need]
    args = []
    for v, s in views:
      args += [_vp(v), s]
    with _on_device(g.device):
      st = lib.ikl_circle_penalty_backward(*args, grad_out.numel(), ctx.slopes[0], ctx.slopes[1], _vp(g), s_g, *[_vp(o) for o in outs],
                                           _stream(g.device))
    _cabi.check(st, "ikl_circle_penalty_backward")
    grads = [None if o is None else _reduce_to(o, t.shape) for o, t in zip(outs, operands)]
    return (*grads, None, None)


def circle_penalty(jx, jy, ox, oy, radius, inside_slope=1.0, outside_slope=0.1):
  assert inside_slope >= 0 and outside_slope >= 0          # penalties.py:10
  src = jx if isinstance(jx, torch.Tensor) else torch.as_tensor(jx)
  dev = _device_for(jx, jy, ox, oy, radius)
  (a, b, c, d, e), _ = _prep(dev, jx, jy, ox, oy, radius)
  out = _CirclePenalty.apply(a, b, c, d, e, float(inside_slope), float(outside_slope))
  if src.device != dev:
    out = out.to(device=src.device)
  return out


class _JointLimitPenalty(torch.autograd.Function):
  @staticmethod
  def forward(ctx, theta, lo, hi, inside_slope, outside_slope):
    lib = _cabi.load()
    shape = theta.shape if theta.shape == lo.shape == hi.shape else torch.broadcast_shapes(theta.shape, lo.shape, hi.shape)
    views = [_flat_view(t, shape) for t in (theta, lo, hi)]
    out = torch.empty(shape, dtype=torch.float32, device=theta.device)
    args = []
    for v, s in views:
      args += [_vp(v), s]
    with _on_device(theta.device):
      st = lib.ikl_joint_limit_penalty_forward(*args, out.numel(), inside_slope, outside_slope, _vp(out), _stream(theta.device))
    _cabi.check(st, "ikl_joint_limit_penalty_forward")
    ctx.save_for_backward(theta, lo, hi)
    ctx.slopes = (inside_slope, outside_slope)
    return out

  @staticmethod
  def backward(ctx, grad_out):
    lib = _cabi.load()
    operands = ctx.saved_tensors
    shape = grad_out.shape
    views = [_flat_view(t, shape) for t in operands]
    g, s_g = _flat_view(grad_out if grad_out.dtype == torch.float32 else grad_out.to(torch.float32), shape)   # trainer.py:204: stride 0
    need = ctx.needs_input_grad[:3]
    outs = [torch.empty(shape, dtype=torch.float32, device=g.device) if n else None for n in need]
    args = []
    for v, s in views:
      args += [_vp(v), s]
    with _on_device(g.device):
      st = lib.ikl_joint_limit_penalty_backward(*args, grad_out.numel(), ctx.slopes[0], ctx.slopes[1], _vp(g), s_g, *[_vp(o) for o in outs],
                                                _stream(g.device))
    _cabi.check(st, "ikl_joint_limit_penalty_backward")
    grads = [None if o is None else _reduce_to(o, t.shape) for o, t in zip(outs, operands)]
    return (*grads, None, None)


def joint_limit_penalty(theta, limit_min, limit_max, inside_slope=0.1, outside_slope=1.0):
  assert inside_slope >= 0 and outside_slope >= 0          # penalties.py:49
  src = theta if isinstance(theta, torch.Tensor) else torch.as_tensor(theta)
  dev = _device_for(theta, limit_min, limit_max)
  (t, lo, hi), _ = _prep(dev, theta, limit_min, limit_max)
  out = _JointLimitPenalty.apply(t, lo, hi, float(inside_slope), float(outside_slope))
  if src.device != dev:
    out = out.to(device=src.device)
  return out
