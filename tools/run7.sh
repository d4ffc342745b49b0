#!/bin/bash
# scratch driver for one gpurun call (round 2): full suite, batched64 A/B, bench with extras, sweep A/B
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -12
B200OT_BATCHED_REG=0 python tools/batched_probe.py 2>&1 | tail -1
python tools/batched_probe.py 2>&1 | tail -1
python bench.py --steps 5 --warmup 3 > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err
python -c "
import json;d=json.load(open('gpurun_out/r2e_bench.json'));print(d['value'],d['roofline']['frac'],d['e2e']['value'],d['clocks'],d['parity']['ok']);print(json.dumps(d['extra'],indent=1))"
tail -3 gpurun_out/r2e_bench.err
