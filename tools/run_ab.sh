#!/bin/bash
# same-box A/B of the single-sweep kernel forms (bench.py, 65536^2, 5 timed steps of 100 iterations each)
run() { python bench.py --steps 5 --warmup 3 --no-extras --no-parity --no-cpu --e2e-steps 1 2>/dev/null | tail -1 | python -c "
import json,sys;d=json.loads(sys.stdin.read());print(round(d['value'],1),round(d['roofline']['frac'],4),d['clocks']['sm_mhz'],d['fused_iteration_launches'])"; }
for i in 1 2 3; do
  [ -f tools/_ab/libb200ot_r1sweep.so ] && { echo -n "r1-lib     "; B200OT_LIB=$PWD/tools/_ab/libb200ot_r1sweep.so run; }
  echo -n "plain      "; run
  echo -n "x2 (div)   "; B200OT_SWEEP_X2=1 run
  echo -n "x2 (inc)   "; B200OT_SWEEP_X2=2 run
  echo -n "FUSE=1     "; B200OT_FUSE=1 run
done
