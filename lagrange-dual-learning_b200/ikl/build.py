"""Build libikl.so (the sm_100a CUDA kernels + C ABI) in-tree with nvcc.

    python lagrange-dual-learning_b200/ikl/build.py [--force] [--variant NAME -DFLAG=...]

The shared object lands in lagrange-dual-learning_b200/lib/ (git-ignored, but it travels to the GPU
box with the gpurun snapshot).  Objects are rebuilt only when a source or header is newer.
nvcc cross-compiles sm_100a without a GPU, so this runs in the CPU-only build container.
"""
import argparse
import concurrent.futures
import hashlib
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_DIR = os.path.join(PKG_DIR, "lib")
INCLUDE = os.path.join(os.path.dirname(PKG_DIR), "include")

ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON_FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
                "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def nvcc_path():
  for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
    if cand and os.path.exists(cand):
      return cand
  raise RuntimeError("nvcc not found; libikl.so cannot be built (there is no CPU fallback)")


def lib_path(variant=""):
  return os.path.join(LIB_DIR, "libikl%s.so" % (("_" + variant) if variant else ""))


def _sources():
  return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers():
  hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h", ".inl"))]
  hs += [os.path.join(INCLUDE, f) for f in os.listdir(INCLUDE) if f.endswith(".h")]
  return hs


def _stamp(defines):
  h = hashlib.sha1()
  for f in _sources() + sorted(_headers()):
    with open(f, "rb") as fh:
      h.update(fh.read())
  h.update(" ".join(COMMON_FLAGS + ARCH_FLAGS + list(defines)).encode())
  return h.hexdigest()


def _compile_one(args):
  nvcc, src, obj, defines = args
  cmd = [nvcc] + ARCH_FLAGS + COMMON_FLAGS + list(defines) + ["-I", INCLUDE, "-c", src, "-o", obj]
  r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
  return src, r.returncode, r.stdout


def is_current(variant="", defines=()):
  """True when libikl[_variant].so exists and was built from exactly the sources, headers and flags in the tree now."""
  out = lib_path(variant)
  stamp_file = out + ".stamp"
  return os.path.exists(out) and os.path.exists(stamp_file) and open(stamp_file).read() == _stamp(defines)


def build(force=False, variant="", defines=(), verbose=False):
  """Compile every csrc/*.cu for sm_100a and link libikl[_variant].so.  Returns its path.  Serialised across processes by
  a file lock (torchrun starts one process per GPU; only the first one compiles)."""
  import fcntl
  os.makedirs(LIB_DIR, exist_ok=True)
  with open(os.path.join(LIB_DIR, ".build.lock"), "w") as lock:
    fcntl.flock(lock, fcntl.LOCK_EX)
    try:
      return _build_locked(force, variant, defines, verbose)
    finally:
      fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(force, variant, defines, verbose):
  out = lib_path(variant)
  stamp_file = out + ".stamp"
  stamp = _stamp(defines)
  if not force and os.path.exists(out) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
    return out
  nvcc = nvcc_path()
  # objects live outside the tree (they are large and must not travel with the gpurun snapshot; only lib/*.so does)
  import tempfile
  base = os.environ.get("IKL_BUILD_DIR") or os.path.join(tempfile.gettempdir(), "ikl_build_" + hashlib.sha1(PKG_DIR.encode()).hexdigest()[:10])
  obj_dir = os.path.join(base, "obj_" + (variant or "default"))
  os.makedirs(obj_dir, exist_ok=True)
  jobs = [(nvcc, s, os.path.join(obj_dir, os.path.basename(s)[:-3] + ".o"), tuple(defines)) for s in _sources()]
  logs = []
  with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
    for src, rc, log in ex.map(_compile_one, jobs):
      logs.append("== %s\n%s" % (os.path.basename(src), log))
      if rc != 0:
        sys.stderr.write(log)
        raise RuntimeError("nvcc failed on %s" % src)
  with open(os.path.join(obj_dir, "ptxas.log"), "w") as f:
    f.write("\n".join(logs))
  if verbose:
    print("\n".join(logs))
  # the static CUDA runtime stays private to the library: only the ikl_* ABI of include/ikl.h is exported
  link = [nvcc] + ARCH_FLAGS + ["-shared", "-o", out] + [j[2] for j in jobs] + ["-cudart", "static", "-Xlinker", "--exclude-libs,ALL"]
  r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
  if r.returncode != 0:
    sys.stderr.write(r.stdout)
    raise RuntimeError("link of %s failed" % out)
  with open(stamp_file, "w") as f:
    f.write(stamp)
  return out


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--force", action="store_true")
  ap.add_argument("--variant", default="")
  ap.add_argument("--verbose", action="store_true")
  ap.add_argument("defines", nargs="*", help="extra -D flags, e.g. -DIKL_PLANES_VEC=2")
  a = ap.parse_args()
  print(build(force=a.force, variant=a.variant, defines=a.defines, verbose=a.verbose))


if __name__ == "__main__":
  main()
