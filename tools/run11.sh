#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fot_cost or feature_coupling or fot or cotl" 2>&1 | tail -8
timeout 300 python tools/fot_probe.py 2>&1 | tail -6
