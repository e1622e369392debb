"""CPU oracle for the FK + constraint-penalty Lagrangian + dual-ascent hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``lagrange-dual-learning_b200/`` may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl
reference`` legs of ``bench.py`` do, and only as the checker / the timed CPU baseline.

This is a from-scratch restatement, in PyTorch eager ops on the CPU, of the arithmetic
the reference performs on this path (the reference *is* PyTorch eager code, so its own
arithmetic library is torch; pinned ``torch==2.2.0`` in the reference's Pipfile.lock,
torch 2.11.0 in this image).  Every function cites the reference lines it follows.
The op order per element is kept identical to the reference so that fp32 results are
bit-identical to it on the same torch build; the functions are dtype-generic, so the
same code evaluated in fp64 is the "tight truth" the CUDA kernels are also compared to.

PARITY PINNING (see tests/test_oracle_pinned.py and tests/golden/):
  * the reference's own known-answer tests (test/test_training_utils.py:7-41) are
    restated in tests/ against this oracle;
  * tests/golden/make_golden.py imports the unmodified reference from /root/reference
    (forward_kinematics.py, penalties.py and the real IkLagrangeDualTrainer) and stores
    its outputs (fp32 + fp64) as fixtures; this oracle must reproduce them bit-for-bit
    in fp32 for per-sample values and to fp32 round-off for batch sums;
  * the shipped multipliers.pth sizes (1/4/7/16/16) pin the constraint layout.
  * tests/test_oracle_vs_live_reference.py executes the reference's own forward_kinematics.py / penalties.py by file
    path (baseline/_ref or /root/reference) on derandomised hypothesis-drawn inputs -- shapes, slopes in either
    order, reversed limits, exact kinks -- and requires bit equality of values and gradients in fp32 and fp64.

The closed-form gradient (``lagrangian_grad_closed_form``) is an independent derivation
 evaluated in numpy fp64; it cross-checks both torch autograd on
this oracle and the CUDA backward kernel.
"""
from __future__ import annotations

import dataclasses
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

MAX_OBSTACLES = 4  # reference hard-wires (B,4,4): scripts/trainer.py:218, kinematics/dataset.py:80


# ----------------------------------------------------------------------------------------------
# Constraint layout  (reference: scripts/trainer.py:71-99)
# ----------------------------------------------------------------------------------------------
@dataclasses.dataclass
class OracleSpec:
  """What a resources/cfg_*.json + the two slope options boil down to for the hot path."""
  num_links: int = 3
  enforce_joint_limits: bool = False
  enforce_obstacles: bool = False
  dynamic: bool = False
  num_obstacles: int = 0
  joint_limits: Tuple[float, float] = (-3.14159, 3.14159)
  static_obstacles: Tuple[Tuple[float, float, float], ...] = ()   # (x, y, radius)
  good_slope: float = 0.1     # scripts/options.py:35  penalty_good_slope
  bad_slope: float = 1.0      # scripts/options.py:34  penalty_bad_slope
  link_lengths: Tuple[float, ...] = (1.0, 1.0, 1.0)   # utils/constants.py:8-10

  @property
  def num_constraints(self) -> int:
    k = 1
    if self.enforce_joint_limits:
      k += self.num_links
    if self.enforce_obstacles:
      k += self.num_links * self.num_obstacles
    return k


def spec_from_json(cfg: dict, num_links: int = 3, good_slope: float = 0.1, bad_slope: float = 1.0) -> OracleSpec:
  """Parse the reference's constraint JSON the way scripts/trainer.py:75-95,209-216 reads it."""
  enforce_obs = cfg["enforce_obstacles"] == True  # noqa: E712  (reference compares with == True)
  dyn = bool(cfg["dynamic_constraints"]["random_obstacles"] == True)  # noqa: E712
  static = tuple((float(o["x"]), float(o["y"]), float(o["radius"])) for o in cfg["static_constraints"]["obstacles"])
  if enforce_obs:
    n_obs = int(cfg["dynamic_constraints"]["random_obstacles_num"]) if dyn else len(static)
  else:
    n_obs = 0
  lo, hi = cfg["static_constraints"]["joint_limits"]
  return OracleSpec(num_links=num_links,
                    enforce_joint_limits=cfg["enforce_joint_limits"] == True,  # noqa: E712
                    enforce_obstacles=enforce_obs, dynamic=dyn and enforce_obs, num_obstacles=n_obs,
                    joint_limits=(float(lo), float(hi)), static_obstacles=static,
                    good_slope=good_slope, bad_slope=bad_slope,
                    link_lengths=tuple([1.0] * num_links))


def constraint_names(spec: OracleSpec) -> List[str]:
  """Human-readable constraint names in multiplier order (scripts/trainer.py:73-95)."""
  names = ["EE"]
  if spec.enforce_joint_limits:
    names += ["JL%d" % (j + 1) for j in range(spec.num_links)]
  if spec.enforce_obstacles:
    tag = "DYN_OB" if spec.dynamic else "STAT_OB"
    for o in range(spec.num_obstacles):
      names += ["%s%d_J%d" % (tag, o + 1, j + 1) for j in range(spec.num_links)]
  assert len(names) == spec.num_constraints
  return names


def initial_multipliers(spec: OracleSpec, initial_lambda: float) -> torch.Tensor:
  """lambda_0 = initial_lambda * ones(K)  (scripts/trainer.py:99)."""
  return initial_lambda * torch.ones(spec.num_constraints)


# ----------------------------------------------------------------------------------------------
# Forward kinematics  (reference: kinematics/forward_kinematics.py:47-83, and :7-22 for 2 links)
# ----------------------------------------------------------------------------------------------
def forward_kinematics(theta: torch.Tensor, link_lengths: Optional[Sequence[float]] = None):
  """Planar serial-arm FK for a (B, J) batch.

  Returns a tuple ordered END EFFECTOR FIRST, i.e. (x_J, ..., x_2, x_1), each (B, 3) holding
  (x, y, cumulative angle) of that link's tip -- the order of forward_kinematics.py:83.
  Arithmetic per element follows :64-81: phi_k = phi_{k-1} + theta_k left to right;
  x_k = x_{k-1} + L_k*cos(phi_k); the angle column is accumulated the same way (:71,76,81).
  """
  assert theta.dim() == 2
  J = theta.shape[1]
  L = [1.0] * J if link_lengths is None else list(link_lengths)
  tips = []
  phi = theta[:, 0]
  px = L[0] * torch.cos(phi)
  py = L[0] * torch.sin(phi)
  tips.append(torch.stack([px, py, phi], dim=1))
  for k in range(1, J):
    phi = phi + theta[:, k]
    px = px + L[k] * torch.cos(phi)
    py = py + L[k] * torch.sin(phi)
    tips.append(torch.stack([px, py, phi], dim=1))
  return tuple(reversed(tips))


def fk_two_link_scalar(theta: Sequence[float]) -> np.ndarray:
  """2-link EE pose in fp64 (forward_kinematics.py:17-22); the only thing pinning J=2."""
  t0, t1 = float(theta[0]), float(theta[1])
  return np.array([np.cos(t0) + np.cos(t0 + t1), np.sin(t0) + np.sin(t0 + t1), t0 + t1])


# ----------------------------------------------------------------------------------------------
# Penalties  (reference: kinematics/penalties.py:5-17 and :44-56)
# ----------------------------------------------------------------------------------------------
def circle_penalty(jx, jy, ox, oy, radius, inside_slope=1.0, outside_slope=0.1):
  """max(-inside*(d-r), -outside*(d-r)) with d = sqrt((jx-ox)^2 + (jy-oy)^2)   (penalties.py:12-17)."""
  assert inside_slope >= 0 and outside_slope >= 0            # penalties.py:10
  gap = torch.sqrt((jx - ox) ** 2 + (jy - oy) ** 2) - radius
  return torch.maximum(-inside_slope * gap, -outside_slope * gap)


def joint_limit_penalty(theta, limit_min, limit_max, inside_slope=0.1, outside_slope=1.0):
  """max(in*|th-mid| - in*0.5*rng, out*|th-mid| - out*0.5*rng)   (penalties.py:51-56).

  ``in*0.5`` is a Python-double product before it touches the tensor, exactly as in the reference.
  """
  assert inside_slope >= 0 and outside_slope >= 0            # penalties.py:49
  mid = (limit_min + limit_max) / 2.0
  rng = torch.abs(limit_max - limit_min)
  dev = torch.abs(theta - mid)
  return torch.maximum(inside_slope * dev - inside_slope * 0.5 * rng,
                       outside_slope * dev - outside_slope * 0.5 * rng)


# ----------------------------------------------------------------------------------------------
# Lagrangian assembly  (reference: scripts/trainer.py:170-247, FK/penalty part of process_batch)
# ----------------------------------------------------------------------------------------------
def obstacle_table(spec: OracleSpec, batch: int, dynamic_obstacles: Optional[torch.Tensor], dtype) -> torch.Tensor:
  """(B, 4, 4) obstacle parameter block as process_batch sees it (scripts/trainer.py:209-226)."""
  if spec.dynamic:
    assert dynamic_obstacles is not None               # trainer.py:211
    return dynamic_obstacles
  table = torch.zeros(batch, MAX_OBSTACLES, 4)        # fp32 on purpose: trainer.py:218 builds it with torch.zeros
  for o, (x, y, r) in enumerate(spec.static_obstacles):
    table[:, o, 0] = x
    table[:, o, 1] = y
    table[:, o, 2] = r
    table[:, o, 3] = -1
  return table


def lagrangian(theta: torch.Tensor, q_ee_desired: torch.Tensor, lam: torch.Tensor, spec: OracleSpec,
               dynamic_obstacles: Optional[torch.Tensor] = None, per_sample: bool = False) -> Dict[str, torch.Tensor]:
  """Everything process_batch computes after the network (scripts/trainer.py:182-247).

  theta (B,J) is the network output.  Returns the reference's ``outputs`` keys (minus
  pred_joint_theta), plus, with per_sample=True, the (B,K) matrix of un-summed terms.
  Constraint order: [EE, JL_1..JL_J, (o=0: j=0..J-1), (o=1: ...)...] where j indexes the FK
  tuple, i.e. j=0 is the end effector (trainer.py:86-88, :231-240).
  """
  B = theta.shape[0]
  dt = theta.dtype
  tips = forward_kinematics(theta, spec.link_lengths)            # trainer.py:182
  q_ee = tips[0]                                                   # trainer.py:183
  out: Dict[str, torch.Tensor] = {"q_ee_actual": q_ee}
  out["joint_angle_l2"] = 0.01 * (0.5 * theta ** 2).mean()         # trainer.py:187-190 (logged only)
  err = 0.5 * (q_ee_desired[:, :2] - q_ee[:, :2]) ** 2             # trainer.py:192
  out["position_err_sq"] = err
  sums = [err.sum().unsqueeze(0)]                                  # trainer.py:195
  cols = [err.sum(dim=1)] if per_sample else None

  if spec.enforce_joint_limits:                                    # trainer.py:198-205
    lim = torch.tensor(spec.joint_limits, dtype=torch.float32).to(dt).unsqueeze(0).expand(B, -1).to(theta.device)  # trainer.py:170
    jl = []
    for j in range(spec.num_links):
      pj = joint_limit_penalty(theta[:, j], lim[:, 0], lim[:, 1],
                               inside_slope=spec.good_slope, outside_slope=spec.bad_slope)
      jl.append(pj.sum())
      if per_sample:
        cols.append(pj)
    sums.append(torch.stack(jl))

  if spec.enforce_obstacles:                                       # trainer.py:209-241
    obst = obstacle_table(spec, B, dynamic_obstacles, dt).to(theta.device)
    ob = []
    for o in range(spec.num_obstacles):
      ox, oy, rad = obst[:, o, 0], obst[:, o, 1], obst[:, o, 2]
      for j in range(spec.num_links):
        po = circle_penalty(tips[j][:, 0], tips[j][:, 1], ox, oy, rad,
                            inside_slope=spec.bad_slope, outside_slope=spec.good_slope)   # trainer.py:237
        ob.append(po.sum())
        if per_sample:
          cols.append(po)
    if ob:
      sums.append(torch.stack(ob))
    else:
      sums.append(torch.zeros(0, dtype=dt, device=theta.device))

  viol = torch.cat(sums)                                           # trainer.py:244
  out["constraint_violations"] = viol
  out["lagrange_loss"] = (lam.to(dt) * viol).sum()                 # trainer.py:247
  if per_sample:
    out["per_sample_terms"] = torch.stack(cols, dim=1)
  return out


def lagrangian_with_grad(theta, q_ee_desired, lam, spec, dynamic_obstacles=None):
  """Forward + autograd backward exactly as train_with_relaxation does (trainer.py:262-264)."""
  th = theta.detach().clone().requires_grad_(True)
  out = lagrangian(th, q_ee_desired, lam, spec, dynamic_obstacles)
  out["lagrange_loss"].backward()
  out = {k: v.detach() for k, v in out.items()}
  out["dtheta"] = th.grad.detach()
  return out


# ----------------------------------------------------------------------------------------------
# Dual ascent  (reference: scripts/trainer.py:267-286)
# ----------------------------------------------------------------------------------------------
def dual_update(lam: torch.Tensor, batch_violations: torch.Tensor, multiplier_lr: float) -> torch.Tensor:
  """lam <- clamp(lam + lr * sum over validation batches of the K-vectors, min=0)  (trainer.py:274-284).

  ``batch_violations`` is the (n_batches, K) stack the reference fills at :274-278.
  """
  lam = lam + multiplier_lr * batch_violations.sum(dim=0)          # trainer.py:280
  return lam.clamp(min=0)                                          # trainer.py:284


# ----------------------------------------------------------------------------------------------
# Independent closed-form gradient in numpy fp64  (not from the reference)
# ----------------------------------------------------------------------------------------------
def lagrangian_grad_closed_form(theta: np.ndarray, q_ee_desired: np.ndarray, weights: np.ndarray, spec: OracleSpec,
                                dynamic_obstacles: Optional[np.ndarray] = None):
  """Returns (g[K] sums, dL/dtheta (B,J)) in fp64 with torch's tie conventions.

  max() ties split the gradient evenly (torch MaximumBackward), sign(0) = 0 for abs.
  At a circle centre (d == 0) torch yields NaN; here the direction term is taken as 0 --
  the deliberate deviation the CUDA kernels also make ("edge semantics").
  """
  th = np.asarray(theta, dtype=np.float64)
  B, J = th.shape
  w = np.asarray(weights, dtype=np.float64)
  a, b = float(spec.good_slope), float(spec.bad_slope)
  L = np.asarray(spec.link_lengths, dtype=np.float64)
  phi = np.cumsum(th, axis=1)
  c, s = np.cos(phi), np.sin(phi)
  px = np.cumsum(L * c, axis=1)            # px[:, m] = x of tip m+1
  py = np.cumsum(L * s, axis=1)
  K = spec.num_constraints
  g = np.zeros(K)
  Gx = np.zeros((B, J))                     # dL/d(px[:, m]), dL/d(py[:, m])
  Gy = np.zeros((B, J))
  dth = np.zeros((B, J))

  ex = np.asarray(q_ee_desired, dtype=np.float64)[:, 0] - px[:, J - 1]
  ey = np.asarray(q_ee_desired, dtype=np.float64)[:, 1] - py[:, J - 1]
  g[0] = (0.5 * ex * ex + 0.5 * ey * ey).sum()
  Gx[:, J - 1] += -w[0] * ex
  Gy[:, J - 1] += -w[0] * ey
  idx = 1

  if spec.enforce_joint_limits:
    lo = np.float64(np.float32(spec.joint_limits[0]))
    hi = np.float64(np.float32(spec.joint_limits[1]))
    mid, h = (lo + hi) / 2.0, 0.5 * abs(hi - lo)
    for j in range(J):
      u = th[:, j] - mid
      vin, vout = a * np.abs(u) - a * h, b * np.abs(u) - b * h
      g[idx] = np.maximum(vin, vout).sum()
      slope = np.where(vin > vout, a, np.where(vin < vout, b, 0.5 * (a + b)))
      dth[:, j] += w[idx] * slope * np.sign(u)
      idx += 1

  if spec.enforce_obstacles:
    if spec.dynamic:
      ob = np.asarray(dynamic_obstacles, dtype=np.float64)
    else:
      ob = np.zeros((B, MAX_OBSTACLES, 4))
      for o, (x, y, r) in enumerate(spec.static_obstacles):
        ob[:, o, :3] = (np.float64(np.float32(x)), np.float64(np.float32(y)), np.float64(np.float32(r)))
    for o in range(spec.num_obstacles):
      for j in range(J):
        m = J - 1 - j                       # FK tuple index j -> tip m+1
        dx, dy = px[:, m] - ob[:, o, 0], py[:, m] - ob[:, o, 1]
        d = np.sqrt(dx * dx + dy * dy)
        t = d - ob[:, o, 2]
        vin, vout = -b * t, -a * t          # inside slope = bad, outside = good (trainer.py:237)
        g[idx] = np.maximum(vin, vout).sum()
        slope = np.where(vin > vout, -b, np.where(vin < vout, -a, -0.5 * (a + b)))
        with np.errstate(divide="ignore", invalid="ignore"):
          inv = np.where(d > 0, 1.0 / d, 0.0)
        Gx[:, m] += w[idx] * slope * dx * inv
        Gy[:, m] += w[idx] * slope * dy * inv
        idx += 1

  # suffix sums over tips, then the chain rule through phi_m
  Sx = np.cumsum(Gx[:, ::-1], axis=1)[:, ::-1]
  Sy = np.cumsum(Gy[:, ::-1], axis=1)[:, ::-1]
  dphi = L * (-s * Sx + c * Sy)
  dth += np.cumsum(dphi[:, ::-1], axis=1)[:, ::-1]
  return g, dth
