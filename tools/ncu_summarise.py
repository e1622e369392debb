#!/usr/bin/env python
"""Turn an .ncu-rep into the small text artefacts committed under profiles/: the details page (CSV) and a trimmed set
of raw metrics (JSON).   python tools/ncu_summarise.py gpurun_out/prof.ncu-rep profiles/r01/name"""
import csv
import io
import json
import subprocess
import sys

KEEP = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active", "sm__cycles_elapsed.max",
        "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum", "l1tex__t_bytes.sum", "lts__t_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed")


def main():
  rep, out = sys.argv[1], sys.argv[2]
  raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
  rows = list(csv.reader(io.StringIO(raw)))
  hdr, units = rows[0], rows[1]
  summary = []
  for r in rows[2:]:
    d = {"kernel": r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else ""}
    for i, h in enumerate(hdr):
      if h in KEEP or ("issue_stalled" in h and h.endswith("per_issue_active.ratio")):
        d[h] = r[i] + (" " + units[i] if units[i] else "")
    summary.append(d)
  with open(out + ".metrics.json", "w") as f:
    json.dump(summary, f, indent=1)
  det = subprocess.run(["ncu", "-i", rep, "--page", "details", "--csv"], capture_output=True, text=True).stdout
  with open(out + ".details.csv", "w") as f:
    f.write(det)
  print("wrote", out + ".metrics.json", out + ".details.csv")


if __name__ == "__main__":
  main()
