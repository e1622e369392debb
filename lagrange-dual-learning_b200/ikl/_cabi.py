"""ctypes binding of libikl.so -- the only way Python reaches the CUDA kernels (include/ikl.h).

There is NO CPU implementation behind these calls: if the shared object is missing and cannot be
built, or a launch fails (no CUDA device), a RuntimeError is raised.
"""
import ctypes
import os
import threading

_c_f = ctypes.c_float
_c_i32 = ctypes.c_int32
_c_i64 = ctypes.c_int64
_c_vp = ctypes.c_void_p

MAX_LINKS = 3
MAX_OBSTACLES = 4
MAX_CONSTRAINTS = 1 + MAX_LINKS + MAX_LINKS * MAX_OBSTACLES
MAX_RANKS = 8
ABI_VERSION = 5

IKL_OK, IKL_ERR_INVALID_ARGUMENT, IKL_ERR_UNSUPPORTED, IKL_ERR_CUDA = 0, 1, 2, 3


class IklSpec(ctypes.Structure):
  _fields_ = [
      ("num_links", _c_i32), ("enforce_joint_limits", _c_i32), ("num_obstacles", _c_i32), ("dynamic_obstacles", _c_i32),
      ("limit_lo", _c_f), ("limit_hi", _c_f), ("good_slope", _c_f), ("bad_slope", _c_f),
      ("link_length", _c_f * MAX_LINKS),
      ("static_obstacles", (_c_f * 3) * MAX_OBSTACLES),
  ]


class IklBatch(ctypes.Structure):
  _fields_ = [
      ("batch", _c_i64),
      ("theta", _c_vp), ("theta_sb", _c_i64), ("theta_sj", _c_i64),
      ("target", _c_vp), ("target_sb", _c_i64), ("target_sc", _c_i64),
      ("obstacles", _c_vp), ("obst_sb", _c_i64), ("obst_so", _c_i64), ("obst_sc", _c_i64),
  ]


class IklWeights(ctypes.Structure):
  _fields_ = [("host", ctypes.POINTER(_c_f)), ("device", _c_vp)]


class IklExchange(ctypes.Structure):
  _fields_ = [("rank", _c_i32), ("world", _c_i32), ("peers", _c_vp * MAX_RANKS)]


class IklDualStep(ctypes.Structure):
  _fields_ = [("lam_in", _c_vp), ("lam_out", _c_vp), ("multiplier_lr", _c_f)]


# name -> (restype, argtypes); mirrors include/ikl.h one to one (tests/test_cabi_symbols.py parses the header)
_SIGNATURES = {
    "ikl_num_constraints": (_c_i32, [ctypes.POINTER(IklSpec)]),
    "ikl_workspace_bytes": (ctypes.c_size_t, [_c_i64]),
    "ikl_lagrangian_forward": (ctypes.c_int, [ctypes.POINTER(IklSpec), ctypes.POINTER(IklBatch), ctypes.POINTER(IklWeights),
                                              _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp]),
    "ikl_lagrangian_forward_backward": (ctypes.c_int, [ctypes.POINTER(IklSpec), ctypes.POINTER(IklBatch), ctypes.POINTER(IklWeights),
                                                       _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_i64, _c_i64, _c_vp]),
    "ikl_exchange_bytes": (ctypes.c_size_t, []),
    "ikl_lagrangian_forward_allreduce": (ctypes.c_int, [ctypes.POINTER(IklSpec), ctypes.POINTER(IklBatch), ctypes.POINTER(IklWeights),
                                                        _c_vp, _c_vp, _c_vp, _c_vp, ctypes.POINTER(IklExchange), _c_vp]),
    "ikl_lagrangian_forward_backward_allreduce": (ctypes.c_int, [ctypes.POINTER(IklSpec), ctypes.POINTER(IklBatch),
                                                                 ctypes.POINTER(IklWeights), _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_i64,
                                                                 _c_i64, ctypes.POINTER(IklExchange), _c_vp]),
    "ikl_lagrangian_forward_dual_step": (ctypes.c_int, [ctypes.POINTER(IklSpec), ctypes.POINTER(IklBatch), ctypes.POINTER(IklWeights),
                                                        _c_vp, _c_vp, _c_vp, _c_vp, ctypes.POINTER(IklExchange),
                                                        ctypes.POINTER(IklDualStep), _c_vp]),
    "ikl_lagrangian_forward_backward_dual_step": (ctypes.c_int, [ctypes.POINTER(IklSpec), ctypes.POINTER(IklBatch),
                                                                 ctypes.POINTER(IklWeights), _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_i64,
                                                                 _c_i64, ctypes.POINTER(IklExchange), ctypes.POINTER(IklDualStep),
                                                                 _c_vp]),
    "ikl_allreduce_bytes": (ctypes.c_size_t, [_c_i64, _c_i32]),
    "ikl_allreduce_sum_f32": (ctypes.c_int, [_c_vp, _c_i64, ctypes.POINTER(IklExchange), _c_i64, _c_vp]),
    "ikl_lagrangian_backward": (ctypes.c_int, [ctypes.POINTER(IklSpec), ctypes.POINTER(IklBatch), ctypes.POINTER(IklWeights),
                                               _c_vp, _c_vp, _c_i64, _c_i64, _c_vp]),
    "ikl_dual_update": (ctypes.c_int, [_c_vp, _c_vp, _c_vp, _c_f, _c_i32, _c_vp, _c_vp]),
    "ikl_fk_forward": (ctypes.c_int, [_c_vp, _c_i64, _c_i64, _c_i64, _c_i32, ctypes.POINTER(_c_f), _c_vp, _c_vp]),
    "ikl_fk_backward": (ctypes.c_int, [_c_vp, _c_i64, _c_i64, _c_i64, _c_i32, ctypes.POINTER(_c_f), _c_vp, _c_vp, _c_vp]),
    "ikl_circle_penalty_forward": (ctypes.c_int, [_c_vp, _c_i64] * 5 + [_c_i64, _c_f, _c_f, _c_vp, _c_vp]),
    "ikl_circle_penalty_backward": (ctypes.c_int, [_c_vp, _c_i64] * 5 + [_c_i64, _c_f, _c_f, _c_vp, _c_i64] + [_c_vp] * 5 + [_c_vp]),
    "ikl_joint_limit_penalty_forward": (ctypes.c_int, [_c_vp, _c_i64] * 3 + [_c_i64, _c_f, _c_f, _c_vp, _c_vp]),
    "ikl_joint_limit_penalty_backward": (ctypes.c_int, [_c_vp, _c_i64] * 3 + [_c_i64, _c_f, _c_f, _c_vp, _c_i64] + [_c_vp] * 3 + [_c_vp]),
    "ikl_dataset_generate": (ctypes.c_int, [ctypes.POINTER(IklSpec), _c_i64, ctypes.c_uint64, _c_f, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp]),
    "ikl_eval_stats": (ctypes.c_int, [ctypes.POINTER(IklSpec), ctypes.POINTER(IklBatch), _c_f, _c_vp, _c_vp, _c_vp]),
    "ikl_set_reserved_sms": (None, [_c_i32]),
    "ikl_abi_version": (_c_i32, []),
    "ikl_strerror": (ctypes.c_char_p, [ctypes.c_int]),
    "ikl_last_cuda_error": (_c_i32, []),
    "ikl_launch_count": (_c_i64, []),
    "ikl_last_kernel_name": (ctypes.c_char_p, []),
}

_lib = None
_lib_path = None
_lock = threading.Lock()


def default_lib_path():
  env = os.environ.get("IKL_LIB")
  if env:
    return env
  from . import build as _build
  return _build.lib_path()


def load(path=None):
  """Load (building first if needed and possible) libikl.so and declare the prototypes."""
  global _lib, _lib_path
  with _lock:
    if _lib is not None and (path is None or path == _lib_path):
      return _lib
    p = path or default_lib_path()
    if path is None and not os.environ.get("IKL_LIB"):
      # the default in-tree library must match the sources in the tree: rebuild when they changed (a no-op when the
      # stamp matches), and never run stale kernels silently
      from . import build as _build
      if not _build.is_current():
        try:
          p = _build.build()
        except Exception as e:  # no nvcc, compile error ...
          raise RuntimeError("libikl.so is missing or older than csrc/ and could not be rebuilt (%s). The IK Lagrangian path "
                             "has no CPU fallback: run `python lagrange-dual-learning_b200/ikl/build.py`." % (e,))
    elif not os.path.exists(p):
      raise RuntimeError("%s does not exist" % p)
    lib = ctypes.CDLL(p)
    for name, (res, args) in _SIGNATURES.items():
      fn = getattr(lib, name)          # AttributeError here = header/library mismatch: fail loudly
      fn.restype = res
      fn.argtypes = args
    if lib.ikl_abi_version() != ABI_VERSION:
      raise RuntimeError("libikl.so ABI version %d != binding %d" % (lib.ikl_abi_version(), ABI_VERSION))
    _lib, _lib_path = lib, p
    return lib


def loaded_path():
  return _lib_path


def check(status, what):
  if status == IKL_OK:
    return
  lib = load()
  msg = lib.ikl_strerror(status).decode()
  if status == IKL_ERR_CUDA:
    msg += " (cudaError %d)" % lib.ikl_last_cuda_error()
  raise RuntimeError("%s failed: %s" % (what, msg))


def set_reserved_sms(n):
  load().ikl_set_reserved_sms(int(n))


def launch_count():
  return int(load().ikl_launch_count())


def last_kernel_name():
  return load().ikl_last_kernel_name().decode()
