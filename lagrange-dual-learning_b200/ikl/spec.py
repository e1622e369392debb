"""Constraint layout of the IK Lagrangian: what a resources/cfg_*.json plus the two slope options
mean for the kernels.  Mirrors scripts/trainer.py:71-99 (constraint count, names, lambda_0) and
:170, :203-204, :237 (limits and slope assignment) of the reference.
"""
import ctypes
import math

import numpy as np

from . import _cabi


class ConstraintSpec(object):
  """Host-side description of one task configuration; owns the IklSpec passed over the C ABI."""

  def __init__(self, num_links=3, enforce_joint_limits=False, joint_limits=(-3.14159, 3.14159), num_obstacles=0,
               dynamic=False, static_obstacles=(), good_slope=0.1, bad_slope=1.0, link_lengths=None):
    # the reference asserts slopes >= 0 at every penalty call (kinematics/penalties.py:10,49)
    assert good_slope >= 0 and bad_slope >= 0
    self.num_links = int(num_links)
    self.enforce_joint_limits = bool(enforce_joint_limits)
    self.joint_limits = (float(joint_limits[0]), float(joint_limits[1]))
    self.num_obstacles = int(num_obstacles)
    self.dynamic = bool(dynamic) and self.num_obstacles > 0
    self.static_obstacles = tuple((float(x), float(y), float(r)) for (x, y, r) in static_obstacles)
    self.good_slope = float(good_slope)
    self.bad_slope = float(bad_slope)
    self.link_lengths = tuple([1.0] * self.num_links if link_lengths is None else [float(v) for v in link_lengths])
    if self.num_obstacles > _cabi.MAX_OBSTACLES:
      raise ValueError("at most %d obstacles are supported (the reference hard-wires (B,4,4))" % _cabi.MAX_OBSTACLES)
    if not self.dynamic and self.num_obstacles != len(self.static_obstacles[:self.num_obstacles]):
      raise ValueError("static configuration needs %d obstacle rows" % self.num_obstacles)
    self._c = None

  # -- construction from the reference's inputs ---------------------------------------------------------------
  @classmethod
  def from_config(cls, cfg, num_links=3, good_slope=0.1, bad_slope=1.0, link_lengths=None):
    """cfg: parsed resources/cfg_*.json; slopes: opt.penalty_good_slope / opt.penalty_bad_slope."""
    enforce_obs = cfg["enforce_obstacles"] == True  # noqa: E712  (the reference compares with == True)
    dynamic = cfg["dynamic_constraints"]["random_obstacles"] == True  # noqa: E712
    static = []
    for ob in cfg["static_constraints"]["obstacles"]:
      assert ob["type"] == "circle"                  # scripts/trainer.py:220
      static.append((ob["x"], ob["y"], ob["radius"]))
    if not enforce_obs:
      n_obs, dynamic = 0, False
    elif dynamic:
      n_obs = int(cfg["dynamic_constraints"]["random_obstacles_num"])
    else:
      n_obs = len(static)
    return cls(num_links=num_links, enforce_joint_limits=cfg["enforce_joint_limits"] == True,  # noqa: E712
               joint_limits=cfg["static_constraints"]["joint_limits"], num_obstacles=n_obs, dynamic=dynamic,
               static_obstacles=static[:n_obs] if not dynamic else (), good_slope=good_slope, bad_slope=bad_slope,
               link_lengths=link_lengths)

  # -- layout --------------------------------------------------------------------------------------------------
  @property
  def num_constraints(self):
    return 1 + (self.num_links if self.enforce_joint_limits else 0) + self.num_links * self.num_obstacles

  @property
  def constraint_names(self):
    """Names in multiplier order (scripts/trainer.py:73-95). `_J1` is the END EFFECTOR tip."""
    names = ["EE"]
    if self.enforce_joint_limits:
      names.extend("JL{}".format(i + 1) for i in range(self.num_links))
    tag = "DYN_OB" if self.dynamic else "STAT_OB"
    for oi in range(self.num_obstacles):
      names.extend("{}{}_J{}".format(tag, oi + 1, ji + 1) for ji in range(self.num_links))
    return names

  def initial_multipliers(self, initial_lambda, device=None):
    """lambda_0 = initial_lambda * ones(K)  (scripts/trainer.py:99)."""
    import torch
    return initial_lambda * torch.ones(self.num_constraints, device=device)

  # -- C view --------------------------------------------------------------------------------------------------
  @property
  def c_struct(self):
    if self._c is None:
      c = _cabi.IklSpec()
      c.num_links = self.num_links
      c.enforce_joint_limits = 1 if self.enforce_joint_limits else 0
      c.num_obstacles = self.num_obstacles
      c.dynamic_obstacles = 1 if self.dynamic else 0
      c.limit_lo, c.limit_hi = self.joint_limits              # ctypes rounds to float32 like torch.Tensor(list) does
      c.good_slope, c.bad_slope = self.good_slope, self.bad_slope
      for m in range(_cabi.MAX_LINKS):
        c.link_length[m] = self.link_lengths[m] if m < self.num_links else 1.0
      for o in range(_cabi.MAX_OBSTACLES):
        row = self.static_obstacles[o] if (not self.dynamic and o < self.num_obstacles) else (0.0, 0.0, 0.0)
        for k in range(3):
          c.static_obstacles[o][k] = row[k]
      self._c = c
    return self._c

  def __repr__(self):
    return "ConstraintSpec(J=%d, K=%d, jl=%s, obstacles=%d %s, slopes good=%g bad=%g)" % (
        self.num_links, self.num_constraints, self.enforce_joint_limits, self.num_obstacles,
        "dynamic" if self.dynamic else "static", self.good_slope, self.bad_slope)
