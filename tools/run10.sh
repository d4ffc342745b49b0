#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/t10_pytest.log
cat gpurun_out/t10_pytest.log
timeout 600 tools/run_ab.sh 2>&1 | tee gpurun_out/t10_ab.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/t10_bench.json 2> gpurun_out/t10_bench.err; tail -c 1500 gpurun_out/t10_bench.json
timeout 300 python tools/apply_probe.py --sizes 16384,65536 --out gpurun_out/t10_apply.json 2>&1 | tail -3
