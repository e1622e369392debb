"""One whole dual-learning train step as a CUDA graph.

At the reference's native batch size (8 samples, scripts/options.py:15) a train step is pure launch latency: network
forward (~20 kernels), the Lagrangian, backward (~40 kernels) and Adam.  Capturing
    theta = model(x); (viol, loss) = FusedIkLagrangian(theta, ...); loss.backward(); [all-reduce]; optimizer.step()
once and replaying it removes the Python and launch overhead at any batch size; the fused kernel is launched through the
C ABI on the capturing stream like any other kernel.  The multipliers are read from a device tensor inside the graph
(IklWeights.device), so a dual-ascent update needs no re-capture: copy the new lambda into `self.lam`.

  grad_only=True    the Lagrangian launch is the gradient-only kernel (a training step consumes nothing but
                    lagrange_loss.backward(), scripts/trainer.py:262-265); step() then returns NaN placeholders.
  planes=True       the static input buffers are structure-of-arrays planes and the network's last layer is evaluated
                    transposed (SimpleNetwork.forward_planes), so the TMA-staged planes kernel is the one in the graph --
                    what `--device_data` batches arrive as.
  torch.distributed the flat all-reduce of the gradients (ikl.dist.all_reduce_gradients) is captured in the graph too; every
                    rank replays the same sequence.  With peer_allreduce= (ikl.exchange.PeerAllReduce) it is ONE kernel
                    over NVLink peer memory whose launch counter lives on the device -- no NCCL node in the graph (the
                    captured NCCL all-reduce hung on this pool's two-GPU box, profiles/r02/pytest_multi_graph_hang.log).
"""
import torch

from . import dist as ikl_dist
from . import fused


class GraphedTrainStep(object):
  def __init__(self, model, optimizer, spec, batch_size, device, warmup=3, planes=False, grad_only=False, peer_allreduce=None):
    if not torch.cuda.is_available():
      raise RuntimeError("CUDA graphs need a CUDA device")
    for group in optimizer.param_groups:
      if not group.get("capturable", False):
        raise ValueError("construct the optimizer with capturable=True (its step counter must live on the device)")
    self.model, self.optimizer, self.spec = model, optimizer, spec
    self.planes = bool(planes)
    self.peer_allreduce = peer_allreduce  # ikl.exchange.PeerAllReduce: the gradient all-reduce as one peer-memory kernel in the graph
    self.grad_only = bool(grad_only)      # the step's launch is the gradient-only kernel; self.viol / self.loss are NaN placeholders
    dev = torch.device(device)
    B, K = int(batch_size), spec.num_constraints        # B: THIS rank's share of the batch
    self.batch_size = B
    self.x = torch.zeros(B, 20, device=dev)
    if self.planes:
      self.q_des = torch.zeros(3, B, device=dev).t()                                       # (B,3) view, one plane per column
      self.obst = torch.zeros(4, 3, B, device=dev).permute(2, 0, 1) if spec.dynamic else None   # (B,4,3) view of planes
    else:
      self.q_des = torch.zeros(B, 3, device=dev)
      self.obst = torch.zeros(B, 4, 4, device=dev) if spec.dynamic else None
    self.lam = torch.zeros(K, device=dev)
    self.viol = None
    self.loss = None
    self.graph = None
    self._capture(warmup)

  def _one_step(self):
    theta = self.model.forward_planes(self.x) if self.planes else self.model(self.x)
    # unit_upstream: loss.backward() is called on the Lagrangian itself (scripts/trainer.py:264), so d(loss)/d(theta) is
    # the kernel's output as written -- no rescaling pass
    viol, loss = fused.FusedIkLagrangian.apply(theta, self.q_des, self.obst, self.lam, self.spec, None, True, self.grad_only)
    self.optimizer.zero_grad(set_to_none=True)
    loss.backward()
    ikl_dist.all_reduce_gradients(self.model, peer=self.peer_allreduce)
    self.optimizer.step()
    return viol, loss

  def _capture(self, warmup):
    # valid dummy inputs for the warm-up iterations (they also create the kernel workspace of this stream)
    self.q_des[:, 0] = 1.0
    if self.obst is not None:
      self.obst[:, :, 0] = 5.0
      self.obst[:, :, 2] = 0.4
    # the warm-up steps move the weights and the Adam state: snapshot both (clones -- state_dict() only aliases the live
    # tensors) and put them back IN PLACE afterwards, so the pointers the graph captured stay valid
    weights = {k: v.clone() for k, v in self.model.state_dict().items()}
    had_state = len(self.optimizer.state) > 0
    saved = {p: {k: (v.clone() if torch.is_tensor(v) else v) for k, v in st.items()} for p, st in self.optimizer.state.items()}
    side = torch.cuda.Stream(self.x.device)
    side.wait_stream(torch.cuda.current_stream(self.x.device))
    with torch.cuda.stream(side):
      for _ in range(warmup):
        self._one_step()
      self.graph = torch.cuda.CUDAGraph()
      self.optimizer.zero_grad(set_to_none=True)
      # under torch.distributed NCCL's watchdog thread polls CUDA events while this thread captures: only this thread's
      # calls may be policed by the capture
      mode = "thread_local" if ikl_dist.is_distributed() else "global"
      with torch.cuda.graph(self.graph, stream=side, capture_error_mode=mode):
        self.viol, self.loss = self._one_step()
    torch.cuda.current_stream(self.x.device).wait_stream(side)
    self.model.load_state_dict(weights)
    for p, st in self.optimizer.state.items():
      for name, val in st.items():
        if not torch.is_tensor(val):
          continue
        if had_state and p in saved and torch.is_tensor(saved[p].get(name)):
          val.copy_(saved[p][name])          # e.g. --load_adam: the loaded moments and step count survive the capture
        else:
          val.zero_()                        # a fresh optimizer: step 0, zero moments
    self.q_des.zero_()
    if self.obst is not None:
      self.obst.zero_()

  def close(self):
    """Drop the captured graph.  Needed only when the graph holds a captured NCCL all-reduce (peer_allreduce=None under
    torch.distributed): torch.distributed.destroy_process_group() blocks for ever while such a graph is alive (measured on
    this pool's two-GPU box, profiles/r02/debug_graph_dist_nccl.log: capture, replays and later eager collectives all work,
    the teardown hangs) -- call close() first."""
    self.graph = None

  def set_multipliers(self, lam):
    self.lam.copy_(lam)

  def step(self, inputs):
    """inputs: the loader's dict (input_tensor, q_ee_desired[, dynamic_obstacles]) for exactly batch_size samples, rows or
    planes.  Returns the (static) constraint_violations and lagrange_loss tensors of this step (this rank's shard)."""
    self.x.copy_(inputs["input_tensor"], non_blocking=True)
    self.q_des.copy_(inputs["q_ee_desired"], non_blocking=True)
    if self.obst is not None:
      src = inputs["dynamic_obstacles"]
      ncol = min(src.shape[2], self.obst.shape[2])       # loaders hand (B,4,4) rows [x,y,r,0] or (B,4,3) plane views
      self.obst[:, :, :ncol].copy_(src[:, :, :ncol], non_blocking=True)
    self.graph.replay()
    return self.viol, self.loss
