#!/bin/bash
# what the driver runs at round end, in one call: smoke, the GPU tests, the default bench line, the reference arm
cd /root/repo; mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/final_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['roofline']['frac'], 'e2e', d['e2e']['value'], d['clocks'], d['gpu_launches'], d['cpu_baseline']['value'] if d['cpu_baseline'] else None)
print(d['parity']['ok'], d['parity']['plan_max_norm_err'], d['extra']['online_c4']['ms_per_iteration'], d['extra']['c3_cohort'].get('iterations_per_s'))
PY
[ "$1" = "ref" ] && timeout 600 python bench.py --impl reference --steps 1 --warmup 1 2>/dev/null | tail -1 | cut -c1-300
