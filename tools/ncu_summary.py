#!/usr/bin/env python
"""Turn an .ncu-rep (ncu --set full) into the short text summary kept under profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/r01_sweep.txt
"""
import csv
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__cluster_size", "launch__block_size",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__cycles_elapsed.avg", "smsp__cycles_active.avg",
]
STALL = "smsp__average_warps_issue_stalled_"


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    kidx = hdr.index("Kernel Name")
    for r in rows[2:]:
        print(f"== {r[kidx][:110]}")
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f"  {w:82s} {r[i]:>18s} {units[i]}")
        stalls = [(float(r[i].replace(",", "")), h) for i, h in enumerate(hdr)
                  if h.startswith(STALL) and h.endswith("_per_issue_active.ratio") and r[i]]
        for v, h in sorted(stalls, reverse=True)[:8]:
            print(f"  stall {h[len(STALL):-len('_per_issue_active.ratio')]:40s} {v:8.3f} warps/issue")


if __name__ == "__main__":
    main(sys.argv[1])
