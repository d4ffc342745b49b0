#!/bin/bash
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
for cfg in "1 2" "0 2" "1 0" "1 1" "0 2" "1 2"; do
  set -- $cfg
  echo "== fuse=$1 lag=$2"
  B200OT_FUSE=$1 B200OT_SHARD_LAG=$2 timeout 600 $TR bench.py --gpus $N --steps 5 --warmup 3 --no-extras --no-parity --no-cpu 2>/dev/null | tail -1 | python -c "
import json,sys;d=json.loads(sys.stdin.read());print(d['value'],d['roofline']['frac'],d['e2e']['value'],d['clocks']['sm_mhz'])"
done
