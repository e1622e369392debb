"""End-to-end path for HOST-resident batches: pinned host tensors in, host gradient + sums out.

This is the call a user of the reference makes when the batch comes off a DataLoader
(scripts/trainer.py:165-166 moves inputs[key] to the device, :264 brings nothing back but the loss):
inputs live in (pinned) host memory in the reference's own row-major shapes, and the result -- the K
constraint sums, the Lagrangian and d(loss)/d(theta) -- is wanted on the host.

The batch is cut into chunks that flow through three streams (H2D copy, fused kernel, D2H copy) with
double-buffered device staging, so PCIe transfers in both directions overlap the kernel.  The K sums are
accumulated on the device in float64 across chunks (IklBatch viol_accum) and read back once.
"""
import torch

from . import fused


class HostLagrangianPipeline(object):
  def __init__(self, spec, chunk=1 << 21, device=None, n_buffers=3):
    if not torch.cuda.is_available():
      raise RuntimeError("HostLagrangianPipeline needs a CUDA device (no CPU fallback)")
    self.spec = spec
    self.chunk = int(chunk)
    self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    self.n_buffers = n_buffers
    self.bufs = None
    self._buf_shapes = None
    dev = self.device
    self.s_in = torch.cuda.Stream(dev)
    self.s_compute = torch.cuda.Stream(dev)
    self.s_out = torch.cuda.Stream(dev)
    self.accum = torch.zeros(spec.num_constraints, dtype=torch.float64, device=dev)
    self.accum_host = torch.empty(spec.num_constraints, dtype=torch.float64).pin_memory()
    self.h2d_bytes = 0
    self.d2h_bytes = 0

  def _staging(self, q_cols, obst_shape):
    """Device staging buffers in the shapes of the host tensors: the reference's (B,3) targets / (B,4,4) obstacle rows, or
    the compact (B,2) / (B,n,3) forms that carry only what the path reads (68 instead of 88 bytes per sample over PCIe)."""
    shapes = (q_cols, obst_shape)
    if self._buf_shapes != shapes:
      J, dev, C = self.spec.num_links, self.device, self.chunk
      self.bufs = []
      for _ in range(self.n_buffers):
        self.bufs.append({
            "theta": torch.empty(C, J, dtype=torch.float32, device=dev),
            "q_des": torch.empty(C, q_cols, dtype=torch.float32, device=dev),
            "obst": torch.empty((C,) + tuple(obst_shape), dtype=torch.float32, device=dev) if self.spec.dynamic else None,
            "dtheta": torch.empty(C, J, dtype=torch.float32, device=dev),
            "ready": torch.cuda.Event(), "computed": torch.cuda.Event(), "drained": torch.cuda.Event(),
        })
      self._buf_shapes = shapes
    return self.bufs

  def run(self, theta, q_ee_desired, obstacles, lam, dtheta_out):
    """All tensors are HOST tensors (pinned for full overlap).  lam: K host floats.  Returns (viol (K,) f64 host, loss float).

    q_ee_desired (B,3) or (B,2); obstacles (B,4,4) rows [x,y,r,pad] as the reference's DataLoader hands them, or compact
    (B,n_obs,3).  dtheta_out (B,J) host tensor receives the gradient.  The call returns after everything has landed on the host.
    """
    spec, C = self.spec, self.chunk
    self._staging(q_ee_desired.shape[1], tuple(obstacles.shape[1:]) if spec.dynamic else ())
    B = theta.shape[0]
    lam_list = [float(v) for v in (lam.tolist() if isinstance(lam, torch.Tensor) else lam)]
    cur = torch.cuda.current_stream(self.device)
    for s in (self.s_in, self.s_compute, self.s_out):
      s.wait_stream(cur)
    with torch.cuda.stream(self.s_compute):
      self.accum.zero_()
    self.h2d_bytes = self.d2h_bytes = 0
    for ci, start in enumerate(range(0, B, C)):
      n = min(C, B - start)
      buf = self.bufs[ci % self.n_buffers]
      with torch.cuda.stream(self.s_in):
        if ci >= self.n_buffers:
          self.s_in.wait_event(buf["drained"])         # the buffer's previous gradient has left the device
        buf["theta"][:n].copy_(theta[start:start + n], non_blocking=True)
        buf["q_des"][:n].copy_(q_ee_desired[start:start + n], non_blocking=True)
        self.h2d_bytes += n * (theta.shape[1] + q_ee_desired.shape[1]) * 4
        if spec.dynamic:
          buf["obst"][:n].copy_(obstacles[start:start + n], non_blocking=True)
          self.h2d_bytes += n * obstacles[0].numel() * 4
        buf["ready"].record(self.s_in)
      with torch.cuda.stream(self.s_compute):
        self.s_compute.wait_event(buf["ready"])
        fused.lagrangian_forward_backward(spec, buf["theta"][:n], buf["q_des"][:n], buf["obst"][:n] if spec.dynamic else None,
                                          lam_list, viol_accum=self.accum, dtheta=buf["dtheta"][:n])
        buf["computed"].record(self.s_compute)
      with torch.cuda.stream(self.s_out):
        self.s_out.wait_event(buf["computed"])
        dtheta_out[start:start + n].copy_(buf["dtheta"][:n], non_blocking=True)
        self.d2h_bytes += n * theta.shape[1] * 4
        buf["drained"].record(self.s_out)
    with torch.cuda.stream(self.s_out):
      self.s_out.wait_stream(self.s_compute)
      self.accum_host.copy_(self.accum, non_blocking=True)
      self.d2h_bytes += self.accum_host.numel() * 8
    cur.wait_stream(self.s_out)
    cur.wait_stream(self.s_in)
    self.s_out.synchronize()
    viol = self.accum_host.clone()
    loss = float((viol * torch.tensor(lam_list, dtype=torch.float64)).sum())
    return viol, loss
