#!/usr/bin/env python
"""Per-region stall breakdown of one kernel from an .ncu-rep captured with --import-source on: the SASS is cut at
block barriers / polling loads and the warp-state samples of every region are summed.

    python tools/ncu_regions.py gpurun_out/prof.ncu-rep
"""
import csv
import re
import subprocess
import sys


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    data = rows[2:]
    tot = sum(int(r[ix["# Samples"]]) for r in data)
    texec = sum(int(r[ix["Instructions Executed"]]) for r in data)
    print(rows[0][1][:100], "samples", tot, "warp-instr", texec)
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    cut = [i for i, r in enumerate(data) if re.search(r"BAR\.SYNC|LD\.E\.(64|128)\.STRONG\.SYS|LDG\.E\.(64|128)\.STRONG\.SYS", r[ix["Source"]])]
    prev = 0
    for b in cut + [len(data) - 1]:
        seg = data[prev:b + 1]
        if not seg:
            continue
        smp = sum(int(r[ix["# Samples"]]) for r in seg)
        ex = max(int(r[ix["Instructions Executed"]]) for r in seg)
        nm = sum(1 for r in seg if "MUFU.EX2" in r[ix["Source"]])
        st = {s: sum(int(r[ix[s]] or 0) for r in seg) for s in stalls}
        top = sorted(st.items(), key=lambda kv: -kv[1])[:4]
        if smp * 200 > tot:
            print(f"[{prev:5d},{b:5d}] {100 * smp / tot:5.1f}%  maxexec {ex:9d} ex2 {nm:3d} end: {data[b][ix['Source']].strip()[:40]:40s}",
                  " ".join(f"{k[6:]}={100 * v / tot:.1f}" for k, v in top))
        prev = b + 1


if __name__ == "__main__":
    main(sys.argv[1])
