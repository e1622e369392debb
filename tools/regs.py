#!/usr/bin/env python
"""Register / spill summary from a build's ptxas.log:  python tools/regs.py [variant] [name-substring ...]"""
import re
import sys
variant = sys.argv[1] if len(sys.argv) > 1 else "default"
pats = sys.argv[2:] or ["ELi3ELi4ELb1ELb1ELi"]
import hashlib, os, tempfile
_pkg = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "lagrange-dual-learning_b200")
_base = os.environ.get("IKL_BUILD_DIR") or os.path.join(tempfile.gettempdir(), "ikl_build_" + hashlib.sha1(_pkg.encode()).hexdigest()[:10])
log = open(os.path.join(_base, "obj_%s" % variant, "ptxas.log")).read()
rx = re.compile(r"Compiling entry function '(\S+)' for 'sm_100a'\n(?:ptxas info\s+: Function properties for \S+\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n)?ptxas info\s+: (Used [^\n]*)")
for m in rx.finditer(log):
  n = m.group(1)
  if any(p in n for p in pats) and "lagrangian" in n:
    print(n[30:64], m.group(5), "| stack", m.group(2), "spill", m.group(3), m.group(4))
