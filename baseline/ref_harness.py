"""Run the reference's OWN, unmodified scripts/trainer.py (from baseline/_ref/, installed by tools/install_ref.py) over this
repo's CUDA drop-ins -- level (i) of the boundary: `kinematics`, `utils` and `models` resolve to
lagrange-dual-learning_b200/, the trainer class comes from the reference file.

Two things the reference needs from its environment and this image lacks are supplied here, outside the reference's code:
  * `tensorboardX` (trainer.py:18)                      -> baseline/shims/tensorboardX
  * a git work tree at the cwd (trainer.py:105-107)     -> `git init` + one empty commit in a temporary directory
"""
import contextlib
import importlib.util
import inspect
import os
import subprocess
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "lagrange-dual-learning_b200")
REF_TREE = os.path.join(ROOT, "baseline", "_ref")
SHIMS = os.path.join(ROOT, "baseline", "shims")


def have_reference_tree():
  return os.path.exists(os.path.join(REF_TREE, "scripts", "trainer.py"))


def _load_by_path(name, path):
  spec = importlib.util.spec_from_file_location(name, path)
  mod = importlib.util.module_from_spec(spec)
  sys.modules[name] = mod
  spec.loader.exec_module(mod)
  return mod


def _environment_shims():
  try:
    import tensorboardX  # noqa: F401
  except ImportError:
    if SHIMS not in sys.path:
      sys.path.append(SHIMS)
  try:
    import git  # noqa: F401
  except Exception:            # GitPython (or the git binary it wraps) missing: the commit hash is logged, nothing else
    stub = types.ModuleType("git")

    class Repo(object):
      def __init__(self, *a, **k):
        self.head = types.SimpleNamespace(object=types.SimpleNamespace(hexsha="unknown"))
    stub.Repo = Repo
    sys.modules["git"] = stub


def load_reference_scripts(dropin=True):
  """(trainer module, options module) executed from baseline/_ref/scripts/*.py by FILE PATH, so that no `scripts`
  package on sys.path (this repo has its own fused mirror) can stand in for them.

  dropin=True   `kinematics`, `utils`, `models` resolve to lagrange-dual-learning_b200/ (the CUDA drop-ins, level (i));
  dropin=False  they resolve to baseline/_ref/ too: the reference end to end, as bench.py --impl reference times it.  Only
                in a process that has not imported this repo's drop-in modules."""
  if not have_reference_tree():
    raise RuntimeError("baseline/_ref is not installed: run `python tools/install_ref.py` in the build container")
  _environment_shims()
  first = PKG if dropin else REF_TREE
  if not dropin:
    assert not any(m in sys.modules for m in ("kinematics", "utils", "models")), "drop-in modules already imported in this process"
    sys.path[:] = [p for p in sys.path if os.path.realpath(p) != os.path.realpath(PKG)]
  if first in sys.path:
    sys.path.remove(first)
  sys.path.insert(0, first)
  if "ref_scripts_trainer" in sys.modules:
    return sys.modules["ref_scripts_trainer"], sys.modules["ref_scripts_options"]
  options_mod = _load_by_path("ref_scripts_options", os.path.join(REF_TREE, "scripts", "options.py"))
  trainer_mod = _load_by_path("ref_scripts_trainer", os.path.join(REF_TREE, "scripts", "trainer.py"))
  return trainer_mod, options_mod


def assert_resolution(trainer_mod, dropin=True):
  """The trainer is the reference's file; everything it calls on the hot path is this repo's drop-in (or, with
  dropin=False, the reference's own module)."""
  ref_real, pkg_real = os.path.realpath(REF_TREE), os.path.realpath(PKG if dropin else REF_TREE)
  src = os.path.realpath(inspect.getsourcefile(trainer_mod.IkLagrangeDualTrainer))
  assert src == os.path.join(ref_real, "scripts", "trainer.py"), src
  for name in ("piecewise_circle_penalty", "piecewise_joint_limit_penalty", "ForwardKinematicsThreeLinkTorch", "IkDataset",
               "SimpleNetwork", "save_multipliers", "get_best_system_device"):
    f = os.path.realpath(inspect.getsourcefile(getattr(trainer_mod, name)))
    assert f.startswith(pkg_real + os.sep), (name, f)


@contextlib.contextmanager
def git_worktree(tmp_path):
  """chdir into a fresh git work tree with one commit (the reference records the commit hash, trainer.py:105-107)."""
  d = os.path.join(str(tmp_path), "worktree")
  os.makedirs(d, exist_ok=True)
  env = dict(os.environ, GIT_AUTHOR_NAME="t", GIT_AUTHOR_EMAIL="t@t", GIT_COMMITTER_NAME="t", GIT_COMMITTER_EMAIL="t@t",
             GIT_CONFIG_GLOBAL="/dev/null", GIT_CONFIG_SYSTEM="/dev/null")
  try:
    subprocess.run(["git", "init", "-q", d], check=True, env=env)
    subprocess.run(["git", "-C", d, "commit", "-q", "--allow-empty", "-m", "init"], check=True, env=env)
  except Exception:            # no git binary: _environment_shims() has replaced GitPython by a stub in that case
    pass
  old = os.getcwd()
  os.chdir(d)
  try:
    yield d
  finally:
    os.chdir(old)


def reference_options(options_mod, argv):
  """The reference's Options class parses sys.argv only (options.py:42-44): go through its parser object."""
  return options_mod.Options().parser.parse_args(argv)
