"""CPU checks of the level-(i) drop-in claim: the reference's own scripts/trainer.py (baseline/_ref, installed by
tools/install_ref.py) imports against this repo's tree, the copy is byte-identical to /root/reference, and the recipe of
INTEGRATION.md section 1 resolves every module the way it says.  The compute side is tests/test_gpu_reference_trainer.py."""
import json
import os
import subprocess
import sys

import pytest

from conftest import PKG, ROOT
from baseline import ref_harness as H

REF = "/root/reference"
needs_ref_tree = pytest.mark.skipif(not H.have_reference_tree(), reason="baseline/_ref not installed (python tools/install_ref.py)")


@needs_ref_tree
def test_installed_reference_is_untouched():
  sys.path.insert(0, os.path.join(ROOT, "tools"))
  import install_ref
  n = install_ref.check()
  assert n >= 40
  if os.path.isdir(REF):                     # build container: the copy equals the mounted reference, file by file
    with open(os.path.join(H.REF_TREE, "MANIFEST.json")) as f:
      manifest = json.load(f)["files"]
    for rel, digest in manifest.items():
      assert install_ref.sha256(os.path.join(REF, rel)) == digest, rel


@needs_ref_tree
def test_reference_trainer_resolves_against_the_dropin_tree():
  trainer_mod, options_mod = H.load_reference_scripts()
  H.assert_resolution(trainer_mod)
  opt = H.reference_options(options_mod, ["--model_name", "m", "--config_file", "c.json"])
  assert opt.batch_size == 8 and opt.penalty_good_slope == 0.1          # scripts/options.py:15,35


@needs_ref_tree
def test_integration_recipe_module_resolution(tmp_path):
  """`cd <reference>/scripts; PYTHONPATH=<repo>/lagrange-dual-learning_b200 python train.py ...`.
  Checked here up to the point where train.py would parse its arguments: `scripts.trainer` / `scripts.options` must be the
  REFERENCE's files and `kinematics` / `utils` / `models` this repo's."""
  probe = (
      "import sys, os, inspect; sys.argv = ['train.py']\n"
      "src = open('train.py').read().split('if __name__')[0]\n"          # the reference's own import block (train.py:1-3)
      "exec(compile(src, 'train.py', 'exec'))\n"
      "import scripts.trainer as t, scripts.options as o, kinematics.penalties as p, kinematics.forward_kinematics as f\n"
      "import utils.training_utils as u, models.simple_network as m, kinematics.dataset as d\n"
      "print('\\n'.join(os.path.realpath(x.__file__) for x in (t, o, p, f, u, m, d)))\n")
  env = dict(os.environ, PYTHONPATH=os.pathsep.join([PKG, H.SHIMS]))
  r = subprocess.run([sys.executable, "-c", probe], cwd=os.path.join(H.REF_TREE, "scripts"), env=env, capture_output=True, text=True)
  assert r.returncode == 0, r.stderr[-2000:]
  paths = r.stdout.strip().splitlines()[-7:]
  ref_real, pkg_real = os.path.realpath(H.REF_TREE), os.path.realpath(PKG)
  assert paths[0] == os.path.join(ref_real, "scripts", "trainer.py")
  assert paths[1] == os.path.join(ref_real, "scripts", "options.py")
  for p in paths[2:]:
    assert p.startswith(pkg_real + os.sep), p
