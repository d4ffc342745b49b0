#!/bin/bash
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -12
bash tools/run_stagger.sh
echo "== unfused"; B200OT_FUSE=0 python bench.py --steps 5 --warmup 3 --no-extras --no-parity --no-cpu --e2e-steps 1 2>/dev/null | tail -1 | python -c "
import json,sys;d=json.loads(sys.stdin.read());print(round(d['value'],1),round(d['roofline']['frac'],4),d['clocks']['sm_mhz'])"
python bench.py --steps 5 --warmup 3 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err
python -c "
import json;d=json.load(open('gpurun_out/r2g_bench.json'));print(d['value'],d['roofline']['frac'],d['e2e']['value'],d['clocks'],d['parity']['ok'],d['fused_iteration_launches'],d['fused_iteration_fallbacks']);print(json.dumps(d['extra'],indent=1)[:6000])"
tail -3 gpurun_out/r2g_bench.err
