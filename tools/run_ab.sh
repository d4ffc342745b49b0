#!/bin/bash
# same-box A/B of the single-sweep kernel forms (bench.py, 65536^2, 5 timed steps of 100 iterations each)
run() { python bench.py --steps 5 --warmup 3 --no-extras --no-parity --no-cpu --e2e-steps 1 2>/dev/null | tail -1 | python -c "
import json,sys;d=json.loads(sys.stdin.read());print(round(d['value'],1),round(d['roofline']['frac'],4),d['clocks']['sm_mhz'],d['fused_iteration_launches'])"; }
for i in 1 2; do
  echo -n "plain      "; run
  echo -n "diag1 noX  "; B200OT_SWEEP_DIAG=1 run
  echo -n "var2 1warp "; B200OT_SWEEP_DIAG=2 run
  echo -n "diag3 noXW "; B200OT_SWEEP_DIAG=3 run
  echo -n "pipe       "; B200OT_FUSED_VARIANT=pipe run
done
