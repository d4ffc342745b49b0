#!/bin/bash
# launch list + full capture of the default single-sweep kernel (each ncu pass only after the same command exited 0)
O=gpurun_out
B="python bench.py --steps 1 --warmup 3 --iters 20 --no-cpu --e2e-steps 1 --no-extras --no-parity"
$B > $O/r3_plain1.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02b_launches_bench_iters20.csv $B > $O/r3_ncu1.log 2>&1
$B --graph 0 > $O/r3_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:sweep_lite_plain_kernel -s 30 -c 2 -o $O/r02b_sweep_lite_plain $B --graph 0 > $O/r3_ncu2.log 2>&1
for f in 1 2; do tail -n 2 $O/r3_ncu$f.log; done
ls -la $O/*.ncu-rep
