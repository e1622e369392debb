"""IkLagrangeDualTrainer on the B200 kernels -- the host-side mirror of the reference's scripts/trainer.py:23-341.

Same class name, constructor argument (the argparse Namespace of scripts/options.py), methods (process_batch,
train_with_relaxation, update_multipliers, validate, main), `outputs` dictionary keys, constraint ordering/names,
log-directory contents and checkpoint layout.  What changes underneath:

  * process_batch evaluates FK + all penalties + the K batch sums + the Lagrangian (and, when gradients are enabled,
    d(Lagrangian)/d(theta) in the same pass) in ONE fused CUDA kernel (ikl.fused.FusedIkLagrangian) instead of ~330
    forward / ~630 backward eager kernels.  `--unfused` keeps the reference's op-by-op structure on the
    function-level CUDA drop-ins (kinematics.forward_kinematics / kinematics.penalties).
  * update_multipliers keeps the running sum of the per-batch K-vectors on the device in float64 and applies
    lambda <- max(0, lambda + lr * sum) with ikl_dual_update (trainer.py:274-284).
  * With torch.distributed initialised (torchrun), every rank processes its shard of each batch; the model gradients
    travel in ONE flat all-reduce per step and the K violation sums of the dual step in one all-reduce of K doubles
    (ikl.dist), so lambda and the weights stay identical on all ranks.

The network itself (models.simple_network.SimpleNetwork) stays on PyTorch / cuBLAS.
"""
import json
import os
import time

import torch
from torch.optim import Adam
from torch.utils.data import DataLoader

from ikl import dist as ikl_dist
from ikl import fused
from ikl.spec import ConstraintSpec
from kinematics.dataset import IkDataset
from kinematics.forward_kinematics import ForwardKinematicsThreeLinkTorch
from kinematics.penalties import piecewise_circle_penalty, piecewise_joint_limit_penalty
from models.simple_network import SimpleNetwork
from utils.training_utils import (count_parameters, get_best_system_device, load_model, load_multipliers, save_model,
                                  save_multipliers)

device = get_best_system_device()


class _NullWriter(object):
  def add_scalar(self, *a, **k): pass
  def add_scalars(self, *a, **k): pass
  def close(self): pass


def _summary_writer(log_path):
  """tensorboardX (the reference's choice) if installed, else torch's bundled writer, else a no-op."""
  try:
    from tensorboardX import SummaryWriter
    return SummaryWriter(logdir=log_path)
  except Exception:
    pass
  try:
    from torch.utils.tensorboard import SummaryWriter
    return SummaryWriter(log_dir=log_path)
  except Exception:
    return _NullWriter()


def _git_sha():
  try:
    import git
    return git.Repo(search_parent_directories=True).head.object.hexsha
  except Exception:
    return "unknown"


class BatchOutputs(dict):
  """The `outputs` dict of process_batch.  Entries the training loop never reads (q_ee_actual, position_err_sq,
  joint_angle_l2) are materialised on first access so the fused kernel does not have to write them every step.  Every way
  of reading a plain dict sees them: indexing, get(), `in`, keys(), iteration, len(), items(), values(), dict(outputs)."""

  def __init__(self, lazy):
    super(BatchOutputs, self).__init__()
    self._lazy = lazy

  def _materialise(self, key=None):
    for k in ([key] if key is not None else list(self._lazy.keys())):
      if k in self._lazy:                    # one thunk may fill several entries (q_ee_actual and position_err_sq)
        self._lazy.pop(k)(self)

  def __missing__(self, key):
    if key in self._lazy:
      self._materialise(key)
      return dict.__getitem__(self, key)
    raise KeyError(key)

  def get(self, key, default=None):
    if key in self._lazy:
      self._materialise(key)
    return dict.get(self, key, default)

  def __contains__(self, key):
    return dict.__contains__(self, key) or key in self._lazy

  def __len__(self):
    return dict.__len__(self) + len(self._lazy)

  def keys(self):
    return list(dict.keys(self)) + list(self._lazy.keys())

  def __iter__(self):
    return iter(self.keys())

  def items(self):
    self._materialise()
    return dict.items(self)

  def values(self):
    self._materialise()
    return dict.values(self)

  def copy(self):
    self._materialise()
    return dict(dict.items(self))


class IkLagrangeDualTrainer(object):
  """Train a network to approximate inverse-kinematics solutions of a three-link planar arm by Lagrangian dual
  learning: Adam epochs on sum_i lambda_i g_i(theta) with lambda fixed, alternating with one dual-ascent step.

  Network input (20 floats): desired x, y; joint limit min, max; 4 obstacles x (x, y, radius, flag).
  """

  def __init__(self, opt, datasets=None):
    torch.backends.cudnn.benchmark = True
    self.opt = opt
    with open(self.opt.config_file, "r") as f:
      self.json_config = json.load(f)

    self.rank, self.world = ikl_dist.rank(), ikl_dist.world_size()
    self.device = torch.device(device if device != "cuda" else "cuda:%d" % torch.cuda.current_device())
    self.model = SimpleNetwork(20, self.opt.num_links, hidden_units=self.opt.hidden_units, depth=8, dropout=0.05).to(self.device)
    ikl_dist.broadcast_parameters(self.model)
    self.use_graph = bool(getattr(self.opt, "cuda_graph", False)) and self.device.type == "cuda"
    # Data-parallel runs sum the replicas' parameter gradients once per step.  On a CUDA box that is ONE kernel over NVLink
    # peer memory (ikl.exchange.PeerAllReduce: no NCCL launch, capturable in the step graph); IKL_PEER_ALLREDUCE=0 keeps
    # the NCCL all-reduce instead.
    self._peer_allreduce = None
    if self.world > 1 and self.device.type == "cuda" and os.environ.get("IKL_PEER_ALLREDUCE", "1") != "0":
      from ikl.exchange import PeerAllReduce
      self._peer_allreduce = PeerAllReduce(sum(p.numel() for p in self.model.parameters()), device=self.device)
    # The dual step closes inside its last forward launch (fused.DualStep: running sums exchanged over NVLink peer memory
    # once per epoch + the lambda update, one kernel); IKL_FUSED_DUAL_STEP=0 keeps the separate all-reduce + ikl_dual_update.
    self._fused_dual_step = self.device.type == "cuda" and os.environ.get("IKL_FUSED_DUAL_STEP", "1") != "0"
    self._peer_exchange = None
    if self._fused_dual_step and self.world > 1:
      if os.environ.get("IKL_PEER_ALLREDUCE", "1") != "0":
        from ikl.exchange import PeerExchange
        self._peer_exchange = PeerExchange(device=self.device)
      else:
        self._fused_dual_step = False
    if self.use_graph and self.world > 1 and self._peer_allreduce is None and os.environ.get("IKL_GRAPH_DISTRIBUTED", "0") != "1":
      # Capturing the flat NCCL all-reduce of the gradients inside the step graph is implemented (ikl/graph_step.py) but hung
      # on the two-GPU box of this pool (profiles/r02/pytest_multi_graph_hang.log): with NCCL as the exchange the graph is
      # off unless IKL_GRAPH_DISTRIBUTED=1.  The default exchange (peer-memory kernel) has no NCCL node and is graphed.
      print("NOTE: --cuda_graph with the NCCL gradient all-reduce is experimental (set IKL_GRAPH_DISTRIBUTED=1); running eagerly")
      self.use_graph = False
    if self.use_graph and self.opt.batch_size % self.world != 0:
      print("NOTE: --cuda_graph needs --batch_size divisible by the world size (every rank replays the same graph); running eagerly")
      self.use_graph = False
    self._unit_upstream = False
    self._grad_only = False                    # inside train_with_relaxation: gradient-only launches (unless --train_sums)
    self.optimizer = Adam(self.model.parameters(), lr=self.opt.optimizer_lr, betas=(0.9, 0.999), capturable=self.use_graph)
    self._graph_step = None

    # constraint layout, names and lambda_0  (reference trainer.py:71-99)
    self.spec = ConstraintSpec.from_config(self.json_config, self.opt.num_links, self.opt.penalty_good_slope,
                                           self.opt.penalty_bad_slope)
    if self.spec.enforce_joint_limits:
      print("NOTE: Enforcing joint limit constraints")
    if self.spec.num_obstacles or self.json_config["enforce_obstacles"] == True:  # noqa: E712
      print("NOTE: Enforcing obstacle constraints ({} {})".format(self.spec.num_obstacles, "dynamic" if self.spec.dynamic else "static"))
    self.constraint_names = self.spec.constraint_names
    self.lambda_multipliers = self.spec.initial_multipliers(self.opt.initial_lambda, device=self.device)
    self._lam_host = None                    # (tensor, version, list): host copy handed to the kernel by value

    self.log_path = os.path.join(self.opt.log_dir, self.opt.model_name)
    if self.rank == 0:
      os.makedirs(self.log_path, exist_ok=True)
      self.opt.commit_has = _git_sha()         # (sic) the reference's attribute name
      with open(os.path.join(self.log_path, "opt.json"), "w") as f:
        f.write(json.dumps(self.opt.__dict__, sort_keys=True, indent=4) + "\n")
      with open(os.path.join(self.log_path, "config.json"), "w") as f:
        f.write(json.dumps(self.json_config, indent=4) + "\n")
      with open(os.path.join(self.log_path, "constraint_names.txt"), "w") as f:
        for i, name in enumerate(self.constraint_names):
          f.write("{},{}\n".format(i, name))
      self.writer = _summary_writer(self.log_path)
      print("=======================================\n")
      print("OPTIONS:")
      for label, value in ((" Model parameters", count_parameters(self.model)), (" Lagrange iters", self.opt.lagrange_iters),
                           (" Train iters", self.opt.train_iters), (" Train dataset size", self.opt.train_dataset_size),
                           (" Validation dataset size", self.opt.val_dataset_size), (" Num workers", self.opt.num_workers),
                           (" Num Lagrange multipliers", len(self.lambda_multipliers)),
                           (" Initial multiplier values", self.opt.initial_lambda), (" Logging to", self.log_path),
                           (" Config file", self.opt.config_file), (" Load weights folder", self.opt.load_weights_folder),
                           (" Load adam state?", self.opt.load_adam), (" World size", self.world)):
        print("{}:\n   {}".format(label, value))
      print("=======================================\n")
    else:
      self.writer = _NullWriter()

    if self.opt.load_weights_folder is not None:
      print("NOTE: Load weights path given, going to load model...")
      load_model(self.model, self.optimizer if self.opt.load_adam else None,
                 os.path.join(self.opt.load_weights_folder, "model.pth"),
                 os.path.join(self.opt.load_weights_folder, "adam.pth"), map_location=self.device)
      if getattr(self.opt, "load_multipliers", False):
        self.lambda_multipliers = load_multipliers(self.opt.load_weights_folder).to(self.device)

    if datasets is not None:
      self.train_dataset, self.val_dataset = datasets
    elif getattr(self.opt, "device_data", False):
      # benchmark-scale path: generate on the GPU (one kernel launch), keep everything device-resident
      self.train_dataset = IkDataset.generate_on_device(self.opt.train_dataset_size, self.opt.num_links, self.json_config, seed=0,
                                                        device=self.device)
      self.val_dataset = IkDataset.generate_on_device(self.opt.val_dataset_size, self.opt.num_links, self.json_config, seed=1,
                                                      device=self.device)
    else:
      train_cache = val_cache = None
      if self.opt.cache_save_path is not None:
        fmt = os.path.join(self.opt.cache_save_path, "{}_{}.pt")
        train_cache, val_cache = fmt.format(self.opt.model_name, "train"), fmt.format(self.opt.model_name, "val")
      self.train_dataset = IkDataset(self.opt.train_dataset_size, self.opt.num_links, self.json_config, cache_save_path=train_cache)
      self.val_dataset = IkDataset(self.opt.val_dataset_size, self.opt.num_links, self.json_config, cache_save_path=val_cache)

    val_batch = getattr(self.opt, "val_batch_size", 0) or self.opt.batch_size
    if getattr(self.opt, "device_data", False):
      from ikl.device_loader import DeviceBatchLoader
      self.train_loader = DeviceBatchLoader(self.train_dataset, self.opt.batch_size, True, device=self.device, seed=11, planes=True)
      self.val_loader = DeviceBatchLoader(self.val_dataset, val_batch, True, device=self.device, seed=12, planes=True)
    else:
      self.train_loader = ikl_dist.make_loader(self.train_dataset, self.opt.batch_size, self.opt.num_workers)
      self.val_loader = ikl_dist.make_loader(self.val_dataset, val_batch, self.opt.num_workers)

  # ---------------------------------------------------------------------------------------------------------------
  def _host_multipliers(self, lamda):
    """K Python floats of `lamda`, cached per tensor version: one device->host read per dual epoch, not per step."""
    c = self._lam_host
    if c is None or c[0] is not lamda or c[1] != lamda._version:
      self._lam_host = (lamda, lamda._version, [float(v) for v in lamda.detach().cpu().tolist()])
    return self._lam_host[2]

  def process_batch(self, inputs, lamda):
    """Network forward, then constraint violations and Lagrangian of the batch (reference trainer.py:159-249)."""
    if getattr(self.opt, "unfused", False):
      return self.process_batch_unfused(inputs, lamda)
    for key in inputs:
      inputs[key] = inputs[key].to(self.device, non_blocking=True)
    # device-resident data arrives as planes: evaluate the last layer transposed so theta is planes too and the
    # TMA-staged planes kernel applies; otherwise the reference's row-major tensors go through the rows kernel
    planes_in = inputs["q_ee_desired"].dim() == 2 and inputs["q_ee_desired"].shape[0] > 1 and inputs["q_ee_desired"].stride(0) == 1
    pred_joint_theta = self.model.forward_planes(inputs["input_tensor"]) if planes_in else self.model(inputs["input_tensor"])
    q_ee_desired = inputs["q_ee_desired"]
    obstacles = inputs["dynamic_obstacles"] if self.spec.dynamic else None
    if self.spec.dynamic:
      assert "dynamic_obstacles" in inputs
    spec = self.spec

    def side_outputs(out):
      with torch.no_grad():
        extra = fused.lagrangian_forward(spec, pred_joint_theta.detach(), q_ee_desired, obstacles, None,
                                         want_q_ee=True, want_err_sq=True)
      dict.__setitem__(out, "q_ee_actual", extra["q_ee_actual"])
      dict.__setitem__(out, "position_err_sq", extra["position_err_sq"])
      out._lazy.pop("q_ee_actual", None)
      out._lazy.pop("position_err_sq", None)

    def angle_l2(out):
      # logged only -- NOT part of the loss, exactly as in the reference (trainer.py:187-190 vs :247)
      dict.__setitem__(out, "joint_angle_l2", 0.01 * (0.5 * pred_joint_theta.detach() ** 2).mean())

    outputs = BatchOutputs({"q_ee_actual": side_outputs, "position_err_sq": side_outputs, "joint_angle_l2": angle_l2})
    outputs["pred_joint_theta"] = pred_joint_theta
    viol, loss = fused.FusedIkLagrangian.apply(pred_joint_theta, q_ee_desired, obstacles, lamda, spec,
                                               self._host_multipliers(lamda), self._unit_upstream, self._grad_only)
    outputs["constraint_violations"] = viol
    outputs["lagrange_loss"] = loss
    return outputs

  def process_batch_unfused(self, inputs, lamda):
    """The reference's op-by-op structure on the function-level CUDA drop-ins (level (i) of the boundary)."""
    outputs = {}
    for key in inputs:
      inputs[key] = inputs[key].to(self.device)
    n = len(inputs["q_ee_desired"])
    limits = torch.Tensor(self.json_config["static_constraints"]["joint_limits"]).unsqueeze(0).expand(n, -1).to(self.device)
    theta = self.model(inputs["input_tensor"])
    outputs["pred_joint_theta"] = theta
    tips = ForwardKinematicsThreeLinkTorch(theta)
    outputs["q_ee_actual"] = tips[0]
    outputs["joint_angle_l2"] = 0.01 * (0.5 * theta ** 2).mean()
    err = 0.5 * (inputs["q_ee_desired"][:, :2] - tips[0][:, :2]) ** 2
    outputs["position_err_sq"] = err
    pieces = [err.sum().unsqueeze(0)]
    if self.spec.enforce_joint_limits:
      pieces.append(torch.stack([
          piecewise_joint_limit_penalty(theta[:, j], limits[:, 0], limits[:, 1], inside_slope=self.opt.penalty_good_slope,
                                        outside_slope=self.opt.penalty_bad_slope).sum() for j in range(self.opt.num_links)]))
    if self.json_config["enforce_obstacles"] == True:  # noqa: E712
      if self.spec.dynamic:
        assert "dynamic_obstacles" in inputs
        table = inputs["dynamic_obstacles"]
      else:
        table = torch.zeros(n, 4, 4)
        for o, (x, y, r) in enumerate(self.spec.static_obstacles):
          table[:, o] = torch.Tensor([x, y, r, -1])
        table = table.to(self.device)
      terms = []
      for o in range(self.spec.num_obstacles):
        for j in range(self.opt.num_links):
          terms.append(piecewise_circle_penalty(tips[j][:, 0], tips[j][:, 1], table[:, o, 0], table[:, o, 1], table[:, o, 2],
                                                inside_slope=self.opt.penalty_bad_slope,
                                                outside_slope=self.opt.penalty_good_slope).sum())
      pieces.append(torch.stack(terms) if terms else torch.zeros(0, device=self.device))
    outputs["constraint_violations"] = torch.cat(pieces).to(self.device)
    outputs["lagrange_loss"] = (lamda * outputs["constraint_violations"]).sum()
    return outputs

  # ---------------------------------------------------------------------------------------------------------------
  def train_with_relaxation(self, lamda):
    """One epoch of Adam steps on the Lagrangian relaxation with the multipliers held fixed (trainer.py:251-265)."""
    sampler = getattr(self.train_loader, "batch_sampler", None)
    if hasattr(sampler, "set_epoch"):           # ShardedBatchSampler: a new permutation every relaxation epoch
      self._relaxation_epoch = getattr(self, "_relaxation_epoch", -1) + 1
      sampler.set_epoch(self._relaxation_epoch)
    if self.use_graph and self._graph_step is None:
      from ikl.graph_step import GraphedTrainStep
      self._graph_step = GraphedTrainStep(self.model, self.optimizer, self.spec, self.opt.batch_size // self.world, self.device,
                                          planes=bool(getattr(self.opt, "device_data", False)),
                                          grad_only=not getattr(self.opt, "train_sums", False), peer_allreduce=self._peer_allreduce)
    if self._graph_step is not None:
      self._graph_step.set_multipliers(lamda)
    for ti, inputs in enumerate(self.train_loader):
      if ti >= self.opt.train_iters:
        break
      # a global batch of exactly batch_size samples deals batch_size / world to every rank: all ranks take the same branch
      if self._graph_step is not None and len(inputs["q_ee_desired"]) == self._graph_step.batch_size:
        self._graph_step.step(inputs)          # the whole step (forward, Lagrangian, backward, all-reduce, Adam) is one graph replay
        continue
      self._unit_upstream = True               # .backward() below starts at the Lagrangian itself: no gradient rescaling pass
      # ... and nothing else of `outputs` is consumed here (trainer.py:262-265): the launch is the gradient-only kernel,
      # constraint_violations / lagrange_loss are NaN placeholders carrying the autograd edge (--train_sums: the full kernel)
      self._grad_only = not getattr(self.opt, "train_sums", False) and not getattr(self.opt, "unfused", False)
      try:
        outputs = self.process_batch(inputs, lamda)
      finally:
        self._unit_upstream = False
        self._grad_only = False
      self.model.zero_grad()
      outputs["lagrange_loss"].backward()
      ikl_dist.all_reduce_gradients(self.model, peer=self._peer_allreduce)
      self.optimizer.step()

  def update_multipliers(self, lamda):
    """One dual-ascent step: lambda <- max(0, lambda + multiplier_lr * sum over the validation set of g)  (trainer.py:267-286)."""
    K = len(lamda)
    with torch.no_grad():
      if getattr(self.opt, "unfused", False):
        total = torch.zeros(K, dtype=torch.float64, device=self.device)
        for inputs in self.val_loader:
          total += self.process_batch(inputs, lamda)["constraint_violations"].double()
      else:
        total = torch.zeros(K, dtype=torch.float64, device=self.device)
        host_lam = self._host_multipliers(lamda)
        # the epoch's LAST batch closes the dual step inside its own launch: the rank's running sums cross NVLink once, the
        # block that holds the global sums moves lambda (every rank sees the same number of batches: ShardedBatchSampler)
        close = fused.DualStep(lamda, self.opt.multiplier_lr, exchange=self._peer_exchange) if self._fused_dual_step else None
        batches = iter(self.val_loader)
        upcoming = next(batches, None)
        closed = False
        while upcoming is not None:
          inputs, upcoming = upcoming, next(batches, None)
          for key in inputs:
            inputs[key] = inputs[key].to(self.device, non_blocking=True)
          planes_in = inputs["q_ee_desired"].shape[0] > 1 and inputs["q_ee_desired"].stride(0) == 1
          # train mode (dropout on), as in the reference
          theta = self.model.forward_planes(inputs["input_tensor"]) if planes_in else self.model(inputs["input_tensor"])
          last = upcoming is None and close is not None
          fused.lagrangian_forward(self.spec, theta, inputs["q_ee_desired"],
                                   inputs["dynamic_obstacles"] if self.spec.dynamic else None, host_lam, viol_accum=total,
                                   close_epoch=close if last else None)
          closed = closed or last
        if closed:
          return close.lam_out
      ikl_dist.all_reduce_sum(total)
      return fused.dual_update(lamda, total, self.opt.multiplier_lr)

  def validate(self, epoch, lamda, train_time, mult_time):
    """Validation pass for logging: mean over batches of the batch sums, like the reference (trainer.py:288-323)."""
    n = len(self.val_loader)
    supervised = torch.zeros(n)
    violations = torch.zeros(n, len(lamda))
    lagrange = torch.zeros(n)
    counts = torch.zeros(n)
    with torch.no_grad():
      for vi, inputs in enumerate(self.val_loader):
        outputs = self.process_batch(inputs, lamda)
        counts[vi] = len(inputs["q_ee_desired"])
        supervised[vi] = outputs["joint_angle_l2"] if counts[vi] > 0 else 0.0
        violations[vi, :] = outputs["constraint_violations"]
        lagrange[vi] = outputs["lagrange_loss"]
    if self.world > 1:
      # every rank saw its share of each validation batch: batch sums add up, the per-batch mean of theta^2 is a weighted mean
      pack = torch.cat([violations.reshape(-1), lagrange, supervised * counts, counts]).to(self.device)
      ikl_dist.all_reduce_sum(pack)
      pack = pack.cpu()
      k = violations.numel()
      violations = pack[:k].view_as(violations)
      lagrange, weighted, counts = pack[k:k + n], pack[k + n:k + 2 * n], pack[k + 2 * n:]
      supervised = weighted / counts.clamp(min=1)
    mean_violation = violations.mean(axis=0)
    if self.rank == 0:
      print("==> Epoch {} (Train Time = {:.3f} sec, Mult Time = {:.3f} sec)\n  Orig Loss={}\n  Constraints={}\n  Lagrange Loss={}\n  Multipliers={}".format(
          epoch, train_time, mult_time, supervised.mean(), mean_violation, lagrange.mean(), lamda))
    self.writer.add_scalar("loss/supervised", supervised.mean(), epoch)
    self.writer.add_scalar("loss/lagrange", lagrange.mean(), epoch)
    self.writer.add_scalars("multipliers", {self.constraint_names[i]: v for (i, v) in enumerate(lamda)}, global_step=epoch)
    self.writer.add_scalars("constraints", {self.constraint_names[i]: v for (i, v) in enumerate(mean_violation)}, global_step=epoch)
    self.writer.close()
    return mean_violation

  def main(self):
    """The training loop: relaxation epoch, dual step, validation, periodic checkpoints (trainer.py:325-341)."""
    for epoch in range(self.opt.lagrange_iters):
      t0 = time.time()
      self.train_with_relaxation(self.lambda_multipliers)
      train_time = time.time() - t0
      self.lambda_multipliers = self.update_multipliers(self.lambda_multipliers)
      mult_time = time.time() - t0 - train_time
      self.validate(epoch, self.lambda_multipliers, train_time, mult_time)
      if epoch % self.opt.model_save_hz == 0 and epoch > 0 and self.rank == 0:
        save_model(self.model, self.optimizer, os.path.join(self.log_path, "models"), epoch)
        save_multipliers(self.lambda_multipliers, os.path.join(self.log_path, "models"), epoch)
