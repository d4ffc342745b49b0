#!/bin/bash
# same-box A/B of the single-sweep kernel forms (bench.py, 65536^2, 5 timed steps): it/s, roofline fraction,
# median SM MHz under load, fused launches counted by the library
run() { python bench.py --steps 5 --warmup 3 --no-extras --no-parity --no-cpu --e2e-steps 1 2>/dev/null | tail -1 | python -c "
import json,sys;d=json.loads(sys.stdin.read());print(round(d['value'],1),round(d['roofline']['frac'],4),d['clocks']['sm_mhz'],d['fused_iteration_launches'])"; }
for i in 1 2 3; do
  echo -n "default (plain kernel + finalize)  "; run
  echo -n "B200OT_FUSE=1 (persistent fused)   "; B200OT_FUSE=1 run
  echo -n "B200OT_FUSED_VARIANT=pipe          "; B200OT_FUSED_VARIANT=pipe run
done
