#!/bin/bash
# launch list of the default bench configuration (ncu pass only after the same command exited 0)
O=gpurun_out
B="python bench.py --steps 1 --warmup 3 --iters 20 --no-cpu --e2e-steps 1 --no-extras --no-parity"
$B > $O/r3_plain1.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02c_launches_bench_iters20.csv $B > $O/r3_ncu1.log 2>&1
tail -n 1 $O/r3_ncu1.log | cut -c1-200
