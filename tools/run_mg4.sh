#!/bin/bash
# N-GPU bench lines of the final build: balanced rows (default) and the even split
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519"
mkdir -p gpurun_out
show() { python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('$1', round(d['value'],1), round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['value'],1), d['clocks']['sm_mhz'], d['gpu_launches'], d['parity']['ok'], d['parity']['g_bit_equal_across_ranks'], d['parity']['g_sha256_16'], d['config'].get('balance'))"; }
timeout 600 $TR bench.py --gpus $N --steps 5 --warmup 3 --no-cpu > gpurun_out/mg4_bench_n$N.json 2>gpurun_out/mg4_err.log; show balanced < gpurun_out/mg4_bench_n$N.json
timeout 600 $TR bench.py --gpus $N --steps 5 --warmup 3 --no-cpu --no-extras --no-balance 2>/dev/null | show even
