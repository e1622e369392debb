"""PyTorch-facing side of the fused IK Lagrangian (level (ii) of the drop-in boundary).

  lagrangian_forward            constraint sums + Lagrangian of a batch          (trainer.py:182-247 under no_grad)
  lagrangian_forward_backward   the same plus d(loss)/d(theta) in the same pass  (trainer.py:182-247 + :264)
  FusedIkLagrangian             torch.autograd.Function wrapping the two
  dual_update                   lambda <- max(0, lambda + lr * sum)              (trainer.py:280-284)
  DualStep                      closes a dual step INSIDE the epoch's last forward launch: running sums exchanged over NVLink
                                peer memory once per epoch + the lambda update, one kernel (trainer.py:267-286)

PyTorch is used for device memory and streams only; all arithmetic happens in libikl.so (sm_100a CUDA),
reached through the C ABI in include/ikl.h.  There is no CPU fallback: CPU tensors raise.
"""
import ctypes
import os

import torch

from . import _cabi
from .spec import ConstraintSpec

_workspaces = {}


def _require_cuda_f32(t, name):
  if not isinstance(t, torch.Tensor):
    raise TypeError("%s must be a torch.Tensor" % name)
  if not t.is_cuda:
    raise RuntimeError("%s is on %s: the fused IK Lagrangian runs on CUDA only (no CPU fallback)" % (name, t.device))
  if t.dtype != torch.float32:
    raise TypeError("%s must be float32, got %s" % (name, t.dtype))
  return t


def _stream_ptr(device):
  return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _workspace(device):
  """One zero-initialised workspace per (device, stream): launches on a stream are ordered, so it is reused."""
  key = (device.index if device.index is not None else torch.cuda.current_device(), torch.cuda.current_stream(device).cuda_stream)
  ws = _workspaces.get(key)
  if ws is None:
    ws = torch.zeros(int(_cabi.load().ikl_workspace_bytes(0)), dtype=torch.uint8, device=device)
    _workspaces[key] = ws
  return ws


class _Weights(object):
  """IklWeights from a tensor/sequence: CUDA tensor -> read on device; anything on the host -> by value."""

  def __init__(self, lam, K, device):
    self.c = _cabi.IklWeights()
    self.keep = None
    if isinstance(lam, torch.Tensor) and lam.is_cuda:
      lam = _require_cuda_f32(lam, "lam")
      if lam.numel() != K:
        raise ValueError("lam has %d entries, the configuration has %d constraints" % (lam.numel(), K))
      self.keep = lam.contiguous()
      self.c.host = None
      self.c.device = self.keep.data_ptr()
    else:
      vals = [float(v) for v in (lam.tolist() if isinstance(lam, torch.Tensor) else lam)]
      if len(vals) != K:
        raise ValueError("lam has %d entries, the configuration has %d constraints" % (len(vals), K))
      self.keep = (ctypes.c_float * K)(*vals)
      self.c.host = ctypes.cast(self.keep, ctypes.POINTER(ctypes.c_float))
      self.c.device = None


def _batch_desc(spec, theta, q_ee_desired, obstacles):
  theta = _require_cuda_f32(theta, "theta")
  q_ee_desired = _require_cuda_f32(q_ee_desired, "q_ee_desired")
  if theta.dim() != 2 or theta.shape[1] != spec.num_links:
    raise ValueError("theta must be (B, %d), got %s" % (spec.num_links, tuple(theta.shape)))
  B = theta.shape[0]
  if q_ee_desired.dim() != 2 or q_ee_desired.shape[0] != B or q_ee_desired.shape[1] < 2:
    raise ValueError("q_ee_desired must be (B, >=2), got %s" % (tuple(q_ee_desired.shape),))
  if q_ee_desired.device != theta.device:
    raise ValueError("theta and q_ee_desired live on different devices")
  d = _cabi.IklBatch()
  d.batch = B
  d.theta, d.theta_sb, d.theta_sj = theta.data_ptr(), theta.stride(0), theta.stride(1)
  d.target, d.target_sb, d.target_sc = q_ee_desired.data_ptr(), q_ee_desired.stride(0), q_ee_desired.stride(1)
  if spec.dynamic:
    if obstacles is None:
      raise ValueError("this configuration has dynamic obstacles: pass the (B, n_obs, >=3) obstacle tensor")  # trainer.py:211
    obstacles = _require_cuda_f32(obstacles, "obstacles")
    if obstacles.dim() != 3 or obstacles.shape[0] != B or obstacles.shape[1] < spec.num_obstacles or obstacles.shape[2] < 3:
      raise ValueError("obstacles must be (B, >=%d, >=3), got %s" % (spec.num_obstacles, tuple(obstacles.shape)))
    d.obstacles = obstacles.data_ptr()
    d.obst_sb, d.obst_so, d.obst_sc = obstacles.stride()
  else:
    d.obstacles, d.obst_sb, d.obst_so, d.obst_sc = None, 0, 0, 0
  return d, (theta, q_ee_desired, obstacles)


def _check_accum(viol_accum, K, device):
  if viol_accum is None:
    return None
  if not (viol_accum.is_cuda and viol_accum.dtype == torch.float64 and viol_accum.is_contiguous() and viol_accum.numel() >= K):
    raise ValueError("viol_accum must be a contiguous CUDA float64 tensor with >= K entries")
  return ctypes.c_void_p(viol_accum.data_ptr())


class DualStep(object):
  """What the launch that closes a dual step needs (include/ikl.h: IklDualStep; scripts/trainer.py:267-286).

  Pass as ``close_epoch=`` to the LAST lagrangian_forward / lagrangian_forward_backward of the epoch, together with the
  ``viol_accum`` the earlier batches added into.  That one launch adds its sums to the rank's running sum, exchanges the K
  running sums over NVLink peer memory (``exchange``: ikl.exchange.PeerExchange; None in a single process), leaves the GLOBAL
  epoch sums in ``viol_accum`` and writes ``lam_out = clamp(lam + multiplier_lr * sum, min=0)`` -- bit-identical on every rank
  and to `dual_update` on the same sums.  ``lam``: (K,) float32 CUDA tensor (not modified unless ``out`` is it); ``lam=None``
  only closes the epoch's sum.  The launch's own viol / loss outputs stay LOCAL to the shard."""

  def __init__(self, lam=None, multiplier_lr=0.0, out=None, exchange=None):
    self.exchange = exchange
    self.c = None
    self.lam_out = None
    if lam is not None:
      self.lam_in = _require_cuda_f32(lam, "lam").contiguous()
      if out is None:
        out = torch.empty_like(self.lam_in)
      else:
        _require_cuda_f32(out, "out")
        if not out.is_contiguous() or out.numel() != self.lam_in.numel() or out.device != self.lam_in.device:
          raise ValueError("out must be a contiguous tensor like lam")
      self.lam_out = out
      self.c = _cabi.IklDualStep()
      self.c.lam_in, self.c.lam_out, self.c.multiplier_lr = self.lam_in.data_ptr(), out.data_ptr(), float(multiplier_lr)

  def _args(self, K, device):
    if self.c is not None and (self.lam_in.numel() != K or self.lam_in.device != device):
      raise ValueError("lam must hold the configuration's %d multipliers on %s" % (K, device))
    return (self.exchange.byref() if self.exchange is not None else None, ctypes.byref(self.c) if self.c is not None else None)


def lagrangian_forward(spec, theta, q_ee_desired, obstacles=None, lam=None, viol_accum=None, want_q_ee=False,
                       want_err_sq=False, exchange=None, close_epoch=None):
  """Constraint sums and Lagrangian of one batch, forward only.

  Returns a dict with the reference's output keys: constraint_violations (K,), lagrange_loss (),
  and optionally q_ee_actual (B,3), position_err_sq (B,2)   (scripts/trainer.py:184,193,244,247).
  ``viol_accum`` (K float64 on device) is incremented by the sums: the running total update_multipliers needs.
  ``exchange`` (ikl.exchange.PeerExchange): this call is one rank's shard; the sums are all-reduced inside the kernel over
  NVLink peer memory and every output holds the GLOBAL value (side outputs are not available in that mode).
  ``close_epoch`` (DualStep): this is the last batch of a dual step -- see DualStep; needs ``viol_accum``.
  """
  lib = _cabi.load()
  K = spec.num_constraints
  desc, keep = _batch_desc(spec, theta, q_ee_desired, obstacles)
  dev = theta.device
  w = _Weights(lam if lam is not None else [0.0] * K, K, dev)
  with torch.cuda.device(dev):
    out = torch.empty(K + 1, dtype=torch.float32, device=dev)
    B = theta.shape[0]
    q_ee = torch.empty(B, 3, dtype=torch.float32, device=dev) if want_q_ee else None
    err = torch.empty(B, 2, dtype=torch.float32, device=dev) if want_err_sq else None
    if close_epoch is not None:
      if exchange is not None or want_q_ee or want_err_sq:
        raise ValueError("close_epoch= carries its own exchange and has no side outputs")
      if viol_accum is None:
        raise ValueError("close_epoch= needs the viol_accum the epoch's earlier batches added into")
      ex_ref, dual_ref = close_epoch._args(K, dev)
      st = lib.ikl_lagrangian_forward_dual_step(ctypes.byref(spec.c_struct), ctypes.byref(desc), ctypes.byref(w.c),
                                                ctypes.c_void_p(_workspace(dev).data_ptr()), ctypes.c_void_p(out.data_ptr()),
                                                ctypes.c_void_p(out.data_ptr() + 4 * K), _check_accum(viol_accum, K, dev),
                                                ex_ref, dual_ref, _stream_ptr(dev))
    elif exchange is not None:
      if want_q_ee or want_err_sq:
        raise ValueError("side outputs are per-shard; request them from a call without exchange=")
      st = lib.ikl_lagrangian_forward_allreduce(ctypes.byref(spec.c_struct), ctypes.byref(desc), ctypes.byref(w.c),
                                                ctypes.c_void_p(_workspace(dev).data_ptr()), ctypes.c_void_p(out.data_ptr()),
                                                ctypes.c_void_p(out.data_ptr() + 4 * K), _check_accum(viol_accum, K, dev),
                                                exchange.byref(), _stream_ptr(dev))
    else:
      st = lib.ikl_lagrangian_forward(ctypes.byref(spec.c_struct), ctypes.byref(desc), ctypes.byref(w.c),
                                      ctypes.c_void_p(_workspace(dev).data_ptr()), ctypes.c_void_p(out.data_ptr()),
                                      ctypes.c_void_p(out.data_ptr() + 4 * K), _check_accum(viol_accum, K, dev),
                                      ctypes.c_void_p(q_ee.data_ptr()) if want_q_ee else None,
                                      ctypes.c_void_p(err.data_ptr()) if want_err_sq else None, _stream_ptr(dev))
  _cabi.check(st, "ikl_lagrangian_forward")
  res = {"constraint_violations": out[:K], "lagrange_loss": out[K]}
  if want_q_ee:
    res["q_ee_actual"] = q_ee
  if want_err_sq:
    res["position_err_sq"] = err
  return res


def lagrangian_forward_backward(spec, theta, q_ee_desired, obstacles=None, lam=None, viol_accum=None, dtheta=None, exchange=None,
                                out=None, close_epoch=None):
  """One pass: (constraint_violations (K,), lagrange_loss (), dtheta (B,J)) with dtheta = d(sum_i lam_i g_i)/d(theta).
  With ``exchange`` (ikl.exchange.PeerExchange) the call is one rank's shard and the sums / loss are the GLOBAL ones,
  all-reduced inside the kernel over NVLink peer memory.  ``out``: optional (K+1,) float32 buffer for [sums | loss].
  ``close_epoch`` (DualStep): this launch closes a dual step -- see DualStep; needs ``viol_accum``; sums / loss stay local."""
  lib = _cabi.load()
  K = spec.num_constraints
  if lam is None:
    raise ValueError("lam (the K multipliers / weights) is required for the gradient")
  desc, keep = _batch_desc(spec, theta, q_ee_desired, obstacles)
  dev = theta.device
  w = _Weights(lam, K, dev)
  with torch.cuda.device(dev):
    if out is None:
      out = torch.empty(K + 1, dtype=torch.float32, device=dev)
    if dtheta is None:
      # planes in (theta.stride(0) == 1, possibly a slice of wider planes) -> planes out; otherwise row-major
      if theta.dim() == 2 and theta.shape[0] > 1 and theta.stride(0) == 1:
        dtheta = torch.empty(theta.shape[1], theta.shape[0], dtype=torch.float32, device=dev).t()
      else:
        dtheta = torch.empty(theta.shape, dtype=torch.float32, device=dev)
    else:
      _require_cuda_f32(dtheta, "dtheta")
      if dtheta.shape != theta.shape:
        raise ValueError("dtheta must have theta's shape")
    args = (ctypes.byref(spec.c_struct), ctypes.byref(desc), ctypes.byref(w.c), ctypes.c_void_p(_workspace(dev).data_ptr()),
            ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(out.data_ptr() + 4 * K), _check_accum(viol_accum, K, dev),
            ctypes.c_void_p(dtheta.data_ptr()), dtheta.stride(0), dtheta.stride(1))
    if close_epoch is not None:
      if exchange is not None:
        raise ValueError("close_epoch= carries its own exchange")
      if viol_accum is None:
        raise ValueError("close_epoch= needs the viol_accum the epoch's earlier batches added into")
      st = lib.ikl_lagrangian_forward_backward_dual_step(*args, *close_epoch._args(K, dev), _stream_ptr(dev))
    elif exchange is not None:
      st = lib.ikl_lagrangian_forward_backward_allreduce(*args, exchange.byref(), _stream_ptr(dev))
    else:
      st = lib.ikl_lagrangian_forward_backward(*args, _stream_ptr(dev))
  _cabi.check(st, "ikl_lagrangian_forward_backward")
  return out[:K], out[K], dtheta


def lagrangian_backward(spec, theta, q_ee_desired, obstacles, weights, dtheta=None):
  """d(sum_i w_i g_i)/d(theta) for arbitrary per-constraint weights (host list or CUDA tensor): the backward pass alone.
  On the staged layouts this is the gradient-only kernel (no running sums, no reduction); its d/dtheta equals
  lagrangian_forward_backward's bit for bit."""
  lib = _cabi.load()
  K = spec.num_constraints
  desc, keep = _batch_desc(spec, theta, q_ee_desired, obstacles)
  dev = theta.device
  w = _Weights(weights, K, dev)
  with torch.cuda.device(dev):
    if dtheta is None:
      if theta.dim() == 2 and theta.shape[0] > 1 and theta.stride(0) == 1:
        dtheta = torch.empty(theta.shape[1], theta.shape[0], dtype=torch.float32, device=dev).t()
      else:
        dtheta = torch.empty(theta.shape, dtype=torch.float32, device=dev)
    st = lib.ikl_lagrangian_backward(ctypes.byref(spec.c_struct), ctypes.byref(desc), ctypes.byref(w.c),
                                     ctypes.c_void_p(_workspace(dev).data_ptr()), ctypes.c_void_p(dtheta.data_ptr()),
                                     dtheta.stride(0), dtheta.stride(1), _stream_ptr(dev))
  _cabi.check(st, "ikl_lagrangian_backward")
  return dtheta


def dual_update(lam, viol_sum, multiplier_lr, out=None):
  """lambda <- clamp(lambda + multiplier_lr * viol_sum, min=0)   (scripts/trainer.py:280-284).

  viol_sum: (K,) float64 running sum (from ``viol_accum``) or float32 sum over the validation batches.
  Returns a NEW tensor unless ``out`` is given (the reference rebinds self.lambda_multipliers, trainer.py:333).
  """
  lib = _cabi.load()
  lam = _require_cuda_f32(lam, "lam").contiguous()
  K = lam.numel()
  if not viol_sum.is_cuda or viol_sum.numel() < K or not viol_sum.is_contiguous():
    raise ValueError("viol_sum must be a contiguous CUDA tensor with K entries")
  if viol_sum.dtype == torch.float64:
    v64, v32 = ctypes.c_void_p(viol_sum.data_ptr()), None
  elif viol_sum.dtype == torch.float32:
    v64, v32 = None, ctypes.c_void_p(viol_sum.data_ptr())
  else:
    raise TypeError("viol_sum must be float64 or float32")
  if out is None:
    out = torch.empty_like(lam)
  with torch.cuda.device(lam.device):
    st = lib.ikl_dual_update(ctypes.c_void_p(lam.data_ptr()), v64, v32, float(multiplier_lr), K,
                             ctypes.c_void_p(out.data_ptr()), _stream_ptr(lam.device))
  _cabi.check(st, "ikl_dual_update")
  return out


_CHECK_UNIT_UPSTREAM = os.environ.get("IKL_CHECK_UNIT_UPSTREAM", "0") == "1"      # tests: verify the promise (costs a device sync)


class FusedIkLagrangian(torch.autograd.Function):
  """(constraint_violations[K], lagrange_loss[]) = f(theta, q_ee_desired, obstacles, lam).

  unit_upstream=True is the caller's promise that backward() starts at lagrange_loss itself with gradient 1 (what
  scripts/trainer.py:264 does); the scaling pass `dtheta * grad_loss` is then skipped.

  grad_only=True (needs unit_upstream) is for a training step, which consumes nothing but lagrange_loss.backward()
  (scripts/trainer.py:262-265): the launch is the GRADIENT-ONLY kernel -- d(loss)/d(theta), no running sums, no reduction --
  and the two returned tensors are NaN placeholders that only carry the autograd edge.

  Differentiable w.r.t. theta only (targets, obstacles, limits and lam carry no gradient in the reference).
  When theta needs a gradient the forward launch already produces d(sum_i lam_i g_i)/d(theta); backward
  scales it by the incoming gradient of lagrange_loss.  A non-trivial gradient flowing into
  constraint_violations re-runs the kernel with weights grad_loss*lam + grad_viol.
  """

  @staticmethod
  def forward(ctx, theta, q_ee_desired, obstacles, lam, spec, lam_host=None, unit_upstream=False, grad_only=False):
    ctx.set_materialize_grads(False)
    ctx.spec = spec
    ctx.unit_upstream = bool(unit_upstream)
    ctx.grad_only = bool(grad_only) and ctx.needs_input_grad[0]
    if grad_only and not unit_upstream:
      raise ValueError("grad_only=True needs unit_upstream=True: without the sums only d(lagrange_loss)/d(theta) itself is available")
    weights = lam_host if lam_host is not None else lam
    if ctx.grad_only:
      dtheta = lagrangian_backward(spec, theta.detach(), q_ee_desired, obstacles if spec.dynamic else None, weights)
      nan = torch.full((spec.num_constraints + 1,), float("nan"), dtype=torch.float32, device=theta.device)
      viol, loss = nan[:-1], nan[-1]
      ctx.save_for_backward(theta, q_ee_desired, obstacles if spec.dynamic else None, lam, dtheta)
    elif ctx.needs_input_grad[0]:
      viol, loss, dtheta = lagrangian_forward_backward(spec, theta.detach(), q_ee_desired, obstacles, weights)
      ctx.save_for_backward(theta, q_ee_desired, obstacles if spec.dynamic else None, lam, dtheta)
    else:
      out = lagrangian_forward(spec, theta.detach(), q_ee_desired, obstacles, weights)
      viol, loss = out["constraint_violations"], out["lagrange_loss"]
    return viol, loss

  @staticmethod
  def backward(ctx, grad_viol, grad_loss):
    theta, q_ee_desired, obstacles, lam, dtheta = ctx.saved_tensors
    if ctx.grad_only:
      if grad_viol is not None:
        raise RuntimeError("grad_only=True: the constraint sums were not computed and cannot be differentiated")
      if grad_loss is None:
        return (None,) * 8
      if _CHECK_UNIT_UPSTREAM:
        assert float(grad_loss) == 1.0, "grad_only=True but lagrange_loss received gradient %r" % float(grad_loss)
      return (dtheta,) + (None,) * 7
    if grad_viol is None:
      if grad_loss is None:
        return (None,) * 8
      if ctx.unit_upstream:
        # the caller differentiates lagrange_loss itself (scripts/trainer.py:264: outputs["lagrange_loss"].backward()), so
        # the incoming gradient is 1 and the kernel's d(loss)/d(theta) is returned as written: no extra pass over (B,J)
        if _CHECK_UNIT_UPSTREAM:
          assert float(grad_loss) == 1.0, "unit_upstream=True but lagrange_loss received gradient %r" % float(grad_loss)
        return (dtheta,) + (None,) * 7
      return (dtheta * grad_loss,) + (None,) * 7
    lam_dev = lam if (isinstance(lam, torch.Tensor) and lam.is_cuda) else torch.as_tensor(lam, dtype=torch.float32, device=theta.device)
    w = grad_viol if grad_loss is None else grad_viol + grad_loss * lam_dev
    g = lagrangian_backward(ctx.spec, theta.detach(), q_ee_desired, obstacles, w.contiguous())
    return (g,) + (None,) * 7


def ik_lagrangian(theta, q_ee_desired, obstacles, lam, spec, lam_host=None, unit_upstream=False, grad_only=False):
  """Functional form of FusedIkLagrangian: returns (constraint_violations, lagrange_loss)."""
  if not isinstance(lam, torch.Tensor):
    lam = torch.as_tensor(lam, dtype=torch.float32)
  return FusedIkLagrangian.apply(theta, q_ee_desired, obstacles, lam, spec, lam_host, unit_upstream, grad_only)
