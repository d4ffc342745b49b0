#!/bin/bash
# 2-GPU check of the default (plain kernel, separate launches) against the opt-in persistent fused form
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
one() {
  B200OT_FUSE=$1 timeout 600 $TR bench.py --gpus $N --steps 5 --warmup 3 --no-extras --no-cpu 2>/dev/null | tail -1 | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('fuse=$1',round(d['value'],1),round(d['roofline']['frac'],4),d['clocks']['sm_mhz'],d['fused_iteration_launches'],d['parity']['ok'],d['parity']['g_bit_equal_across_ranks'],d['parity']['g_sha256_16'])"
}
for i in 1 2; do one 0; one 1; done
timeout 900 $TR bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/mg3_bench_n$N.json 2>/dev/null; tail -c 600 gpurun_out/mg3_bench_n$N.json
