#!/bin/bash
# ncu cannot launch the cooperative CLUSTER kernel (driver reports LaunchFailed under the profiler), so the
# single-sweep kernel is profiled in its plain form (B200OT_FUSE=0: same sweep code, finalize as its own launch)
set -x
O=gpurun_out
export B200OT_FUSE=0
python bench.py --steps 1 --warmup 3 --iters 20 --no-cpu --e2e-steps 1 --no-extras --no-parity > $O/r2_plain5.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches_bench_iters20.csv python bench.py --steps 1 --warmup 3 --iters 20 --no-cpu --e2e-steps 1 --no-extras --no-parity > $O/r2_ncu5.log 2>&1
python bench.py --steps 1 --warmup 3 --iters 20 --no-cpu --e2e-steps 1 --no-extras --no-parity --graph 0 > $O/r2_plain6.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:sweep_lite_kernel -s 30 -c 2 -o $O/r02_sweep_lite python bench.py --steps 1 --warmup 3 --iters 20 --no-cpu --e2e-steps 1 --no-extras --no-parity --graph 0 > $O/r2_ncu6.log 2>&1
unset B200OT_FUSE
python tools/batched_probe.py > $O/r2_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:sinkhorn_batched64 -s 9 -c 1 -o $O/r02_batched64 python tools/batched_probe.py > $O/r2_ncu2.log 2>&1
for f in 2 5 6; do tail -n 2 $O/r2_ncu$f.log; done
