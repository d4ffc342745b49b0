#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_fusion_autograd.py -q -x 2>&1 | tail -12
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6
bash tools/run_ncu.sh 2>&1 | tail -30
