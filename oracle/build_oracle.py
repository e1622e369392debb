"""Build the C restatement of the oracle (oracle/ik_oracle.c) with gcc and bind it with ctypes.  Checker only: see the
header of oracle/ik_oracle.py for who may import this."""
import ctypes
import os
import shutil
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libik_oracle.so")
SRC = os.path.join(HERE, "ik_oracle.c")


class OracleSpecC(ctypes.Structure):
  _fields_ = [("enforce_joint_limits", ctypes.c_int), ("num_obstacles", ctypes.c_int), ("dynamic", ctypes.c_int),
              ("limit_lo", ctypes.c_float), ("limit_hi", ctypes.c_float), ("good_slope", ctypes.c_double),
              ("bad_slope", ctypes.c_double), ("static_obstacles", (ctypes.c_float * 3) * 4)]


def build(force=False):
  os.makedirs(OUT_DIR, exist_ok=True)
  if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
    return LIB
  gcc = shutil.which("gcc") or shutil.which("cc")
  if gcc is None:
    raise RuntimeError("no C compiler for the oracle")
  subprocess.run([gcc, "-O2", "-fopenmp", "-shared", "-fPIC", "-o", LIB, SRC, "-lm"], check=True)
  return LIB


_lib = None


def lib():
  global _lib
  if _lib is None:
    _lib = ctypes.CDLL(build())
    _lib.ik_oracle_num_constraints.argtypes = [ctypes.POINTER(OracleSpecC)]
    _lib.ik_oracle_lagrangian_f64.argtypes = [ctypes.POINTER(OracleSpecC), ctypes.c_longlong] + [ctypes.c_void_p] * 7
    _lib.ik_oracle_lagrangian_f32.argtypes = [ctypes.POINTER(OracleSpecC), ctypes.c_longlong] + [ctypes.c_void_p] * 6
    _lib.ik_oracle_diagnostics_f64.argtypes = [ctypes.POINTER(OracleSpecC), ctypes.c_longlong] + [ctypes.c_void_p] * 8
  return _lib


def c_spec(ospec):
  """oracle.ik_oracle.OracleSpec -> OracleSpecC (3 links)."""
  assert ospec.num_links == 3
  s = OracleSpecC()
  s.enforce_joint_limits = 1 if ospec.enforce_joint_limits else 0
  s.num_obstacles = ospec.num_obstacles if ospec.enforce_obstacles else 0
  s.dynamic = 1 if ospec.dynamic else 0
  s.limit_lo, s.limit_hi = ospec.joint_limits
  s.good_slope, s.bad_slope = ospec.good_slope, ospec.bad_slope
  for o, row in enumerate(ospec.static_obstacles[:4]):
    for k in range(3):
      s.static_obstacles[o][k] = row[k]
  return s


def _f32(a):
  return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def lagrangian_f64(ospec, theta, q_des, obst, weights, want_grad=True):
  """Returns (sums[K], abs_sums[K], dtheta (B,3) or None) in float64.  theta/q_des/obst: row-major float32 arrays."""
  L, s = lib(), c_spec(ospec)
  K = L.ik_oracle_num_constraints(ctypes.byref(s))
  th, qd = _f32(theta), _f32(q_des)
  ob = _f32(obst) if (ospec.dynamic and obst is not None) else None
  B = th.shape[0]
  w = np.ascontiguousarray(np.asarray(weights, dtype=np.float64))
  sums, abs_sums = np.zeros(K), np.zeros(K)
  dth = np.zeros((B, 3)) if want_grad else None
  rc = L.ik_oracle_lagrangian_f64(ctypes.byref(s), B, th.ctypes.data, qd.ctypes.data, ob.ctypes.data if ob is not None else None,
                                  w.ctypes.data, sums.ctypes.data, abs_sums.ctypes.data, dth.ctypes.data if want_grad else None)
  assert rc == 0
  return sums, abs_sums, dth


def lagrangian_f32(ospec, theta, q_des, obst, weights, want_grad=True):
  """float arithmetic per sample, float64 sums.  Returns (sums[K] float64, dtheta (B,3) float32 or None)."""
  L, s = lib(), c_spec(ospec)
  K = L.ik_oracle_num_constraints(ctypes.byref(s))
  th, qd = _f32(theta), _f32(q_des)
  ob = _f32(obst) if (ospec.dynamic and obst is not None) else None
  B = th.shape[0]
  w = _f32(weights)
  sums = np.zeros(K)
  dth = np.zeros((B, 3), dtype=np.float32) if want_grad else None
  rc = L.ik_oracle_lagrangian_f32(ctypes.byref(s), B, th.ctypes.data, qd.ctypes.data, ob.ctypes.data if ob is not None else None,
                                  w.ctypes.data, sums.ctypes.data, dth.ctypes.data if want_grad else None)
  assert rc == 0
  return sums, dth


def diagnostics_f64(ospec, theta, q_des, obst, weights, want_terms=False):
  """Per-sample float64 diagnostics (see ik_oracle_diagnostics_f64): dict with kink_dist (B,), cond (B,), tips (B,6) and,
  if requested, terms (B,K)."""
  L, s = lib(), c_spec(ospec)
  K = L.ik_oracle_num_constraints(ctypes.byref(s))
  th, qd = _f32(theta), _f32(q_des)
  ob = _f32(obst) if (ospec.dynamic and obst is not None) else None
  B = th.shape[0]
  w = np.ascontiguousarray(np.asarray(weights, dtype=np.float64))
  kink, cond, tips = np.zeros(B), np.zeros(B), np.zeros((B, 6))
  terms = np.zeros((B, K)) if want_terms else None
  rc = L.ik_oracle_diagnostics_f64(ctypes.byref(s), B, th.ctypes.data, qd.ctypes.data, ob.ctypes.data if ob is not None else None,
                                   w.ctypes.data, kink.ctypes.data, cond.ctypes.data, tips.ctypes.data,
                                   terms.ctypes.data if want_terms else None)
  assert rc == 0
  return {"kink_dist": kink, "cond": cond, "tips": tips, "terms": terms}


def max_threads():
  return int(lib().ik_oracle_max_threads())
