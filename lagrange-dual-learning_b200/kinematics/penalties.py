"""Drop-in for the reference's kinematics/penalties.py (same names / signatures / assertions).

piecewise_circle_penalty and piecewise_joint_limit_penalty run as CUDA kernels (libikl.so) and reproduce
the reference's fp32 results bit for bit; the collision predicates used by dataset generation stay scalar
host logic exactly as in the reference (kinematics/penalties.py:59-92).
"""
import math

import torch

from ikl import ops as _ops


def piecewise_circle_penalty(jx, jy, ox, oy, radius, inside_slope=1.0, outside_slope=0.1) -> torch.Tensor:
  """max(-inside_slope*(d - radius), -outside_slope*(d - radius)), d = distance from (jx, jy) to (ox, oy)."""
  assert(inside_slope >= 0 and outside_slope >= 0)
  return _ops.circle_penalty(jx, jy, ox, oy, radius, inside_slope, outside_slope)


def piecewise_obstacle_penalty(jx, jy, ox, oy, ow, oh, inside_slope=1.0, outside_slope=0.1) -> torch.Tensor:
  """Rectangle penalty: a stub in the reference too (kinematics/penalties.py:27)."""
  assert(inside_slope >= 0 and outside_slope >= 0)
  raise NotImplementedError()


def piecewise_joint_limit_penalty(theta, limit_min, limit_max, inside_slope=0.1, outside_slope=1.0):
  """max(in*|theta-mid| - in*0.5*range, out*|theta-mid| - out*0.5*range) for limits (limit_min, limit_max)."""
  assert(inside_slope >= 0 and outside_slope >= 0)
  return _ops.joint_limit_penalty(theta, limit_min, limit_max, inside_slope, outside_slope)


def no_joint_collisions_rectangle(q_all_joints, x, y, json):
  """True iff no joint tip and not the arm base lies inside the axis-aligned rectangle at (x, y)."""
  assert("width" in json and "height" in json)
  x_hi, y_hi = x + json["width"], y + json["height"]
  points = [(q[0], q[1]) for q in q_all_joints] + [(0, 0)]
  return not any(px >= x and px <= x_hi and py >= y and py <= y_hi for (px, py) in points)


def no_joint_collisions_circle(q_all_joints, x, y, json):
  """True iff every joint tip (x, y, angle rows) and the arm base are strictly outside the circle."""
  assert("radius" in json)
  radius = json["radius"]
  for q in q_all_joints:
    if torch.sqrt((q[0] - x)**2 + (q[1] - y)**2) <= radius:
      return False
  return not (math.sqrt(x**2 + y**2) <= radius)
