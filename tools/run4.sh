#!/bin/bash
# scratch driver for one gpurun call (round 2): fused iteration tests, apply probe, full suite, bench A/B
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "fused_iteration or recovers or paths_match or small_eps or cuda_graph" 2>&1 | tail -8
timeout 600 python tools/apply_probe.py --simt-max 0 --out gpurun_out/r2d_apply_probe.json 2>&1 | tail -4
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err
python -c "
import json;d=json.load(open('gpurun_out/r2d_bench.json'));print(d['value'],d['roofline']['frac'],d['e2e']['value'],d['clocks'],d['gpu_launches'],d['parity']['ok'])"
tail -3 gpurun_out/r2d_bench.err
B200OT_FUSE=0 python bench.py --steps 5 --warmup 3 --no-cpu --no-parity 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('unfused',d['value'],d['roofline']['frac'],d['clocks'])"
