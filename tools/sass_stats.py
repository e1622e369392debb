#!/usr/bin/env python
"""Opcode histogram of one kernel's SASS (cuobjdump), with a rough split into loop body vs rest.

  python tools/sass_stats.py <obj-or-so> <substring of mangled kernel name> [--dump out.sass]
"""
import collections
import re
import subprocess
import sys


def main():
  path, pat = sys.argv[1], sys.argv[2]
  txt = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
  blocks = re.split(r"\n\s+Function : ", txt)
  sel = [b for b in blocks[1:] if pat in b.split("\n", 1)[0]]
  if not sel:
    print("no kernel matches", pat)
    return
  body = sel[0]
  print("kernel:", body.split("\n", 1)[0])
  if "--dump" in sys.argv:
    open(sys.argv[sys.argv.index("--dump") + 1], "w").write(body)
  ins = re.findall(r"/\*([0-9a-f]{4})\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)", body)
  hist = collections.Counter(op for _, op, _ in ins)
  print("total instructions:", len(ins))
  for op, n in hist.most_common(45):
    print("  %-10s %d" % (op, n))


if __name__ == "__main__":
  main()
