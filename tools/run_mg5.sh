#!/bin/bash
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521"
timeout 400 $TR tools/peer_probe.py 65536 200 2>&1 | grep -v "^W\|^\[W\|warn" | tail -2
nvidia-smi --query-gpu=index,clocks.sm,power.draw,temperature.gpu --format=csv,noheader | head -8
