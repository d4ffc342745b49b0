#!/bin/bash
# N-GPU check of the peer loop: merged tail (default) vs separate reduce_push / finalize launches vs persistent fused
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
one() {
  env $1 timeout 600 $TR bench.py --gpus $N --steps 5 --warmup 3 --no-extras --no-cpu 2>/dev/null | tail -1 | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('$1',round(d['value'],1),round(d['roofline']['frac'],4),d['clocks']['sm_mhz'],d['gpu_launches'],d['parity']['ok'],d['parity']['g_bit_equal_across_ranks'],d['parity']['g_sha256_16'])"
}
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fused_iteration or sharded" 2>&1 | tail -3
for i in 1 2; do one B200OT_PEER_TAIL=1; one B200OT_PEER_TAIL=0; done
one B200OT_FUSE=1
timeout 900 $TR bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/mg3_bench_n$N.json 2>/dev/null; tail -c 300 gpurun_out/mg3_bench_n$N.json
