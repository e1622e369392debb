"""CPU-side checks of the C ABI: the library builds, loads, exports every symbol include/ikl.h declares,
its structs match the ctypes mirror, and compute entry points fail loudly without a GPU."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT
from ikl import _cabi, build
from ikl.spec import ConstraintSpec
from ikl.configs import standard_config, STANDARD_CONFIGS


def _declared_functions():
  text = open(os.path.join(ROOT, "include", "ikl.h")).read()
  text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
  return sorted(set(re.findall(r"\b(ikl_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol():
  path = build.build()
  assert os.path.exists(path)
  lib = ctypes.CDLL(path)
  declared = _declared_functions()
  assert len(declared) >= 16
  for name in declared:
    assert hasattr(lib, name), "libikl.so does not export %s" % name
  assert sorted(_cabi._SIGNATURES.keys()) == declared, "ctypes binding and include/ikl.h disagree"


def test_abi_version_and_housekeeping():
  lib = _cabi.load()
  assert lib.ikl_abi_version() == _cabi.ABI_VERSION
  assert lib.ikl_strerror(0) == b"ok"
  assert b"CUDA" in lib.ikl_strerror(3)
  assert lib.ikl_workspace_bytes(1 << 24) >= 2048 * 16 * 8
  assert lib.ikl_launch_count() >= 0


@pytest.mark.parametrize("name,K", list(zip(STANDARD_CONFIGS, (1, 4, 7, 16, 16))))
def test_constraint_count_matches_shipped_multiplier_sizes(name, K):
  spec = ConstraintSpec.from_config(standard_config(name), 3, 0.005, 1.0)
  assert spec.num_constraints == K
  assert _cabi.load().ikl_num_constraints(ctypes.byref(spec.c_struct)) == K
  assert len(spec.constraint_names) == K and spec.constraint_names[0] == "EE"


def test_constraint_names_follow_reference_order():
  spec = ConstraintSpec.from_config(standard_config("cfg_joint_limits_and_4obs_dynamic"), 3)
  assert spec.constraint_names[:5] == ["EE", "JL1", "JL2", "JL3", "DYN_OB1_J1"]
  assert spec.constraint_names[-1] == "DYN_OB4_J3"
  spec = ConstraintSpec.from_config(standard_config("cfg_joint_limits_and_1obs_static"), 3)
  assert spec.constraint_names == ["EE", "JL1", "JL2", "JL3", "STAT_OB1_J1", "STAT_OB1_J2", "STAT_OB1_J3"]


def test_negative_slope_asserts_like_the_reference():
  with pytest.raises(AssertionError):
    ConstraintSpec.from_config(standard_config("cfg_joint_limits"), 3, good_slope=-0.1)


def test_invalid_arguments_are_reported_not_crashed():
  lib = _cabi.load()
  spec = ConstraintSpec.from_config(standard_config("cfg_joint_limits"), 3).c_struct
  assert lib.ikl_lagrangian_forward(ctypes.byref(spec), None, None, None, None, None, None, None, None, None) == _cabi.IKL_ERR_INVALID_ARGUMENT
  assert lib.ikl_dual_update(None, None, None, 0.1, 4, None, None) == _cabi.IKL_ERR_INVALID_ARGUMENT
  assert lib.ikl_fk_forward(None, 3, 1, 4, 5, None, None, None) == _cabi.IKL_ERR_UNSUPPORTED
  assert lib.ikl_circle_penalty_forward(None, 0, None, 0, None, 0, None, 0, None, 0, 4, -1.0, 0.1, None, None) == _cabi.IKL_ERR_INVALID_ARGUMENT


@pytest.mark.skipif(torch.cuda.is_available(), reason="this is the no-GPU behaviour")
def test_product_path_fails_loudly_without_cuda():
  from ikl import fused, ops
  import kinematics.forward_kinematics as fk
  import kinematics.penalties as pen
  spec = ConstraintSpec.from_config(standard_config("cfg_joint_limits"), 3)
  with pytest.raises(RuntimeError):
    fused.lagrangian_forward(spec, torch.zeros(4, 3), torch.zeros(4, 3), None, [1.0] * 4)
  with pytest.raises(RuntimeError):
    fk.ForwardKinematicsThreeLinkTorch(torch.zeros(4, 3))
  with pytest.raises(RuntimeError):
    pen.piecewise_circle_penalty(torch.zeros(3), torch.zeros(3), torch.zeros(1), torch.zeros(1), torch.ones(1))
  # a raw launch without a device reports a CUDA error through the status code
  lib = _cabi.load()
  buf = (ctypes.c_float * 16)()
  st = lib.ikl_dual_update(ctypes.cast(buf, ctypes.c_void_p), None, ctypes.cast(buf, ctypes.c_void_p), 0.1, 4,
                           ctypes.cast(buf, ctypes.c_void_p), None)
  assert st == _cabi.IKL_ERR_CUDA and lib.ikl_last_cuda_error() != 0


def test_reference_trainer_imports_against_dropin_tree_when_mounted():
  """With the drop-in tree first on sys.path, every name scripts/trainer.py and kinematics/dataset.py import resolves."""
  import kinematics.forward_kinematics as fk
  import kinematics.penalties as pen
  for n in ("ForwardKinematicsThreeLinkTorch", "ForwardKinematicsEightLinkTorch", "ForwardKinematicsTwoLink", "ForwardKinematicsThreeLink"):
    assert callable(getattr(fk, n))
  for n in ("piecewise_joint_limit_penalty", "piecewise_circle_penalty", "no_joint_collisions_circle", "no_joint_collisions_rectangle"):
    assert callable(getattr(pen, n))


def test_missing_library_fails_loudly(monkeypatch):
  """No silent fallback: pointing the binding at a library that is not there raises."""
  monkeypatch.setenv("IKL_LIB", "/nonexistent/libikl.so")
  monkeypatch.setattr(_cabi, "_lib", None)
  monkeypatch.setattr(_cabi, "_lib_path", None)
  with pytest.raises(RuntimeError):
    _cabi.load()


def test_product_tree_never_imports_the_oracle():
  """oracle/ is test infrastructure: nothing under the package may reference it."""
  pkg = os.path.join(ROOT, "lagrange-dual-learning_b200")
  for base, _, files in os.walk(pkg):
    if os.sep + "build" in base:
      continue
    for f in files:
      if f.endswith((".py", ".cu", ".cuh", ".h", ".inl")):
        text = open(os.path.join(base, f), errors="ignore").read()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), os.path.join(base, f)
        assert "ik_oracle" not in text, os.path.join(base, f)


def _kernel_sass(mangled):
  import shutil
  import subprocess
  if shutil.which("cuobjdump") is None:
    pytest.skip("cuobjdump not on PATH")
  path = build.build()
  out = subprocess.run(["cuobjdump", "-sass", "-fun", mangled, path], capture_output=True, text=True).stdout
  assert "Function : " + mangled in out, "kernel %s is not in libikl.so" % mangled
  return out


def test_benchmark_kernels_are_sm100a_tma_and_packed_fp32():
  """The kernels bench.py times are what DESIGN.md section 3 / 4.1 says they are: sm_100a code whose loads are TMA
  (tensor-map UTMALDG for planes, bulk UBLKCP for rows) completing on mbarriers (SYNCS), with the per-pair maths on the
  packed fp32 pipe (FFMA2 / FADD2 / FMUL2) and one MUFU.RSQ per tip-circle term.  No GPU needed: SASS of the built library."""
  cfg5 = "NS_3CfgILi%dELi3ELi4ELb1ELb1ELi1ELb1EEEEEvNS_16LagrangianParamsE"     # 3 links, 4 dynamic obstacles, limits, fwd+bwd host lambda
  planes = _kernel_sass("_ZN3ikl28lagrangian_planes_tma_kernelI" + cfg5 % 2)
  rows = _kernel_sass("_ZN3ikl26lagrangian_rows_tma_kernelI" + cfg5 % 1)
  for sass in (planes, rows):
    assert "arch = sm_100a" in sass
    assert sass.count("FFMA2") > 100 and sass.count("FADD2") > 20 and sass.count("FMUL2") > 20
    assert sass.count("MUFU.RSQ") >= 24                     # 12 terms x 2 samples (the exact redo adds its own)
    assert "SYNCS.ARRIVE.TRANS64" in sass and "SYNCS.PHASECHK.TRANS64.TRYWAIT" in sass
    assert not re.search(r"\b(HMMA|UTCHMMA|IMMA)\b", sass)   # not a contraction: no tensor-core instructions, by design
  assert "UTMALDG.2D" in planes and "UTMALDG.3D" in planes
  assert "UBLKCP.S.G" in rows
