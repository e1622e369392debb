"""Constraint configurations of the five published tasks, built programmatically.

The reference reads these from resources/cfg_*.json (schema: SURVEY §5 "config / flags";
e.g. resources/cfg_joint_limits_and_1obs_static.json:1-25).  A user's own JSON files load
unchanged through ``load_config``; the builders below exist so tests, the benchmark and the
smoke run have the five standard tasks without shipping data files.
"""
import json

TWO_PI_OVER_3 = 2.094395      # the literal the reference's configs use for the +-120 degree limits
PI_LITERAL = 3.14159          # cfg_no_constraints.json's (unenforced) limits
STANDARD_RADIUS = 0.4

_STATIC_CENTRES = {
    "cfg_no_constraints": [],
    "cfg_joint_limits": [],
    "cfg_joint_limits_and_1obs_static": [(0.6, 0.6)],
    "cfg_joint_limits_and_4obs_static": [(0.7, 0.7), (-0.7, 0.7), (0.7, -0.7), (-0.7, -0.7)],
    "cfg_joint_limits_and_4obs_dynamic": [],
}
STANDARD_CONFIGS = tuple(_STATIC_CENTRES.keys())


def circle(x, y, radius=STANDARD_RADIUS):
  return {"type": "circle", "radius": radius, "x": x, "y": y}


def make_config(joint_limits, static_circles=(), num_dynamic=0, enforce_joint_limits=True, enforce_obstacles=True):
  """Assemble a config dict with the reference's schema."""
  return {
      "static_constraints": {
          "joint_limits": [joint_limits[0], joint_limits[1]],
          "obstacles": [circle(*c) for c in static_circles],
      },
      "dynamic_constraints": {
          "random_obstacles": num_dynamic > 0,
          "random_obstacles_num": num_dynamic,
          "random_obstacles_template": circle(1.0, 1.0),
      },
      "enforce_joint_limits": enforce_joint_limits,
      "enforce_obstacles": enforce_obstacles,
  }


def standard_config(name):
  """One of the five task configs by its reference file stem (with or without '.json')."""
  name = name[:-5] if name.endswith(".json") else name
  if name not in _STATIC_CENTRES:
    raise KeyError("unknown standard config %r; have %s" % (name, ", ".join(STANDARD_CONFIGS)))
  if name == "cfg_no_constraints":
    return make_config((-PI_LITERAL, PI_LITERAL), enforce_joint_limits=False, enforce_obstacles=False)
  if name == "cfg_joint_limits":
    return make_config((-TWO_PI_OVER_3, TWO_PI_OVER_3), enforce_obstacles=False)
  if name == "cfg_joint_limits_and_4obs_dynamic":
    return make_config((-TWO_PI_OVER_3, TWO_PI_OVER_3), num_dynamic=4)
  return make_config((-TWO_PI_OVER_3, TWO_PI_OVER_3), static_circles=_STATIC_CENTRES[name])


def load_config(path):
  with open(path, "r") as f:
    return json.load(f)


def write_config(cfg, path):
  with open(path, "w") as f:
    json.dump(cfg, f, indent=2)
    f.write("\n")
  return path
