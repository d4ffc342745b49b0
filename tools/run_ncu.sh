#!/bin/bash
# scratch driver: ncu captures of round 2 (one gpurun call, 1 GPU).  Every profiled command first runs plain.
set -x
O=gpurun_out
python tools/online_probe.py 65536 2 > $O/r2_online_plain.json 2>&1 ; cat $O/r2_online_plain.json | tail -1
python tools/online_probe.py 65536 2 $((48<<20)) 2>&1 | tail -1
# 1. apply_tc forward + transposed at 16384
python tools/apply_probe.py --sizes 16384 --simt-max 0 --out $O/r2_probe16k.json > $O/r2_plain1.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:apply_tc_kernel -s 11 -c 3 -o $O/r02_apply_tc python tools/apply_probe.py --sizes 16384 --simt-max 0 --out $O/r2_probe16k_ncu.json > $O/r2_ncu1.log 2>&1
# 2. batched64
python tools/batched_probe.py > $O/r2_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:sinkhorn_batched64 -s 4 -c 1 -o $O/r02_batched64 python tools/batched_probe.py > $O/r2_ncu2.log 2>&1
# 3. egw
python tools/egw_probe.py > $O/r2_plain3.log 2>&1 && \
ncu --set full --clock-control none -k regex:egw_kernel -s 2 -c 1 -o $O/r02_egw python tools/egw_probe.py > $O/r2_ncu3.log 2>&1
# 4. online iteration: launch list
python tools/online_probe.py 65536 1 > $O/r2_plain4.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/r02_online_launches.csv python tools/online_probe.py 65536 1 > $O/r2_ncu4.log 2>&1
# 5. bench launch list
python bench.py --steps 1 --warmup 3 --iters 20 --no-cpu --e2e-steps 1 --no-extras --no-parity > $O/r2_plain5.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches_bench_iters20.csv python bench.py --steps 1 --warmup 3 --iters 20 --no-cpu --e2e-steps 1 --no-extras --no-parity > $O/r2_ncu5.log 2>&1
# 6. fused sweep, full set
python bench.py --steps 1 --warmup 3 --iters 20 --no-cpu --e2e-steps 1 --no-extras --no-parity --graph 0 > $O/r2_plain6.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:sweep_lite_kernel -s 2 -c 1 -o $O/r02_sweep_fused python bench.py --steps 1 --warmup 3 --iters 20 --no-cpu --e2e-steps 1 --no-extras --no-parity --graph 0 > $O/r2_ncu6.log 2>&1
tail -2 $O/r2_ncu*.log
