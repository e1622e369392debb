#!/usr/bin/env python
"""Install the UNMODIFIED reference into baseline/_ref/ (git-ignored; it travels to the GPU box with the gpurun snapshot).

    python tools/install_ref.py [--src /root/reference] [--check]

The reference (miloknowles/lagrange-dual-learning) is a plain Python tree with no build step, so `pip install` has nothing
to build: installing it means copying its text files -- *.py, the resources/cfg_*.json task configs, the experiment
shell scripts and the shipped multipliers -- byte for byte.  A MANIFEST.json with the sha256 of every file is written so
tests/test_reference_dropin.py can verify on the GPU box that what runs there is what /root/reference holds here.

Who uses baseline/_ref:
  * tests/test_reference_dropin.py  -- runs the reference's own scripts/trainer.py over the CUDA drop-ins (level (i));
  * bench.py --impl reference       -- times the reference's own FK / penalty functions on the host cores.
Nothing under lagrange-dual-learning_b200/ reads it.
"""
import argparse
import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEST = os.path.join(ROOT, "baseline", "_ref")
KEEP_EXT = (".py", ".json", ".sh", ".md", ".sdf", ".lock")
KEEP_NAMES = ("Pipfile", "multipliers.pth")
MAX_BYTES = 2 << 20


def wanted(path):
  name = os.path.basename(path)
  return (name.endswith(KEEP_EXT) or name in KEEP_NAMES) and os.path.getsize(path) <= MAX_BYTES


def sha256(path):
  h = hashlib.sha256()
  with open(path, "rb") as f:
    h.update(f.read())
  return h.hexdigest()


def install(src, dest=DEST):
  if not os.path.isdir(src):
    raise SystemExit("reference tree %s not found (it only exists in the build container)" % src)
  if os.path.isdir(dest):
    shutil.rmtree(dest)
  manifest = {}
  for base, dirs, files in os.walk(src):
    dirs[:] = sorted(d for d in dirs if d not in (".git", "__pycache__"))
    for f in sorted(files):
      p = os.path.join(base, f)
      if not wanted(p):
        continue
      rel = os.path.relpath(p, src)
      out = os.path.join(dest, rel)
      os.makedirs(os.path.dirname(out), exist_ok=True)
      shutil.copyfile(p, out)
      manifest[rel] = sha256(out)
  with open(os.path.join(dest, "MANIFEST.json"), "w") as f:
    json.dump({"source": src, "files": manifest}, f, indent=1, sort_keys=True)
  return manifest


def check(dest=DEST):
  """Every installed file still has the recorded hash (nothing edited the copy).  Returns the number of files."""
  with open(os.path.join(dest, "MANIFEST.json")) as f:
    manifest = json.load(f)["files"]
  for rel, digest in manifest.items():
    if sha256(os.path.join(dest, rel)) != digest:
      raise SystemExit("baseline/_ref/%s differs from the installed reference" % rel)
  return len(manifest)


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--src", default="/root/reference")
  ap.add_argument("--check", action="store_true")
  a = ap.parse_args()
  if a.check:
    print("baseline/_ref: %d files match MANIFEST.json" % check())
    return
  m = install(a.src)
  print("installed %d reference files into %s" % (len(m), DEST))


if __name__ == "__main__":
  main()
