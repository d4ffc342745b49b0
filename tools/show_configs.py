#!/usr/bin/env python
"""Print the size sweep of a tools/bench_configs.py JSON in one line per (size, variant)."""
import json
import sys

j = json.load(open(sys.argv[1]))
for k, v in j.get("sinkhorn_200it_by_size", {}).items():
    for lab, r in v.items():
        print(k, lab, "%.2f us/it  %.0f it/s  %.0f GB/s" % (r["us_per_iteration"], r["iterations_per_s"], r["matrix_GBps"]))
print(json.dumps(j.get("feature_coupling_pot")))
print("C3 200 it ms", j["C3_cohort_4096x4096_d512"]["sinkhorn_200it_ms"], "C1", j["C1_64x64_200it_numpy_api"]["b200_ms"])
