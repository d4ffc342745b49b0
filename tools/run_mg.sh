#!/bin/bash
# scratch driver for a multi-GPU gpurun call: tools/run_mg.sh N [size]
N=${1:-2}
SZ=${2:-32768}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
echo "== peer_probe fused"; timeout 600 $TR tools/peer_probe.py $SZ 200 2>&1 | grep -v "^W\|^\[W\|warn" | tail -3
echo "== peer_probe unfused"; B200OT_FUSE=0 timeout 600 $TR tools/peer_probe.py $SZ 200 2>&1 | grep -v "^W\|^\[W\|warn" | tail -3
echo "== bench fused"; timeout 900 $TR bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2f_bench_n$N.json 2> gpurun_out/r2f_bench_n$N.err
python -c "
import json
d=json.loads(open('gpurun_out/r2f_bench_n$N.json').read().strip().splitlines()[-1])
print(d['value'],d['roofline']['frac'],d['e2e']['value'],d['clocks'],d['parity']);print(json.dumps(d['extra']))"
tail -3 gpurun_out/r2f_bench_n$N.err
echo "== bench unfused"; B200OT_FUSE=0 timeout 900 $TR bench.py --gpus $N --steps 5 --warmup 3 --no-extras --no-parity 2>/dev/null | tail -1 | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('unfused',d['value'],d['roofline']['frac'])"
