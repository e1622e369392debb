"""One whole dual-learning train step as a CUDA graph.

At the reference's native batch size (8 samples, scripts/options.py:15) a train step is pure launch latency: network
forward (~20 kernels), the Lagrangian, backward (~40 kernels) and Adam.  Capturing
    theta = model(x); (viol, loss) = FusedIkLagrangian(theta, ...); loss.backward(); optimizer.step()
once and replaying it removes the Python and launch overhead; the fused kernel is launched through the C ABI on the
capturing stream like any other kernel.  The multipliers are read from a device tensor inside the graph (IklWeights.device),
so a dual-ascent update needs no re-capture: copy the new lambda into `self.lam`.
"""
import torch

from . import fused


class GraphedTrainStep(object):
  def __init__(self, model, optimizer, spec, batch_size, device, warmup=3):
    if not torch.cuda.is_available():
      raise RuntimeError("CUDA graphs need a CUDA device")
    for group in optimizer.param_groups:
      if not group.get("capturable", False):
        raise ValueError("construct the optimizer with capturable=True (its step counter must live on the device)")
    self.model, self.optimizer, self.spec = model, optimizer, spec
    dev = torch.device(device)
    B, K = batch_size, spec.num_constraints
    self.x = torch.zeros(B, 20, device=dev)
    self.q_des = torch.zeros(B, 3, device=dev)
    self.obst = torch.zeros(B, 4, 4, device=dev) if spec.dynamic else None
    self.lam = torch.zeros(K, device=dev)
    self.viol = None
    self.loss = None
    self.graph = None
    self._capture(warmup)

  def _one_step(self):
    theta = self.model(self.x)
    viol, loss = fused.FusedIkLagrangian.apply(theta, self.q_des, self.obst, self.lam, self.spec)
    self.optimizer.zero_grad(set_to_none=True)
    loss.backward()
    self.optimizer.step()
    return viol, loss

  def _capture(self, warmup):
    # valid dummy inputs for the warm-up iterations (they also create the kernel workspace of this stream)
    self.q_des[:, 0] = 1.0
    if self.obst is not None:
      self.obst[:, :, 0] = 5.0
      self.obst[:, :, 2] = 0.4
    state = ({k: v.clone() for k, v in self.model.state_dict().items()}, self.optimizer.state_dict())
    side = torch.cuda.Stream(self.x.device)
    side.wait_stream(torch.cuda.current_stream(self.x.device))
    with torch.cuda.stream(side):
      for _ in range(warmup):
        self._one_step()
      self.graph = torch.cuda.CUDAGraph()
      self.optimizer.zero_grad(set_to_none=True)
      with torch.cuda.graph(self.graph, stream=side):
        self.viol, self.loss = self._one_step()
    torch.cuda.current_stream(self.x.device).wait_stream(side)
    # the warm-up steps moved the weights and the Adam moments: put the caller's state back
    self.model.load_state_dict(state[0])
    fresh = self.optimizer.state_dict()
    for k, st in fresh["state"].items():
      for name, val in st.items():
        if torch.is_tensor(val):
          val.zero_()
    self.q_des.zero_()
    if self.obst is not None:
      self.obst.zero_()

  def set_multipliers(self, lam):
    self.lam.copy_(lam)

  def step(self, inputs):
    """inputs: the DataLoader's dict (input_tensor, q_ee_desired[, dynamic_obstacles]) for exactly batch_size samples.
    Returns the (static) constraint_violations and lagrange_loss tensors of this step."""
    self.x.copy_(inputs["input_tensor"], non_blocking=True)
    self.q_des.copy_(inputs["q_ee_desired"], non_blocking=True)
    if self.obst is not None:
      self.obst.copy_(inputs["dynamic_obstacles"], non_blocking=True)
    self.graph.replay()
    return self.viol, self.loss
