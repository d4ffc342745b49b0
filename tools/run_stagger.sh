#!/bin/bash
# scratch: A/B the cluster start stagger of the single-sweep kernel on one box
for st in 0 350 700 1000 1400 0 700; do
  echo -n "stagger=$st  "
  B200OT_STAGGER=$st python bench.py --steps 5 --warmup 3 --no-extras --no-parity --no-cpu --e2e-steps 1 2>/dev/null | tail -1 | python -c "
import json,sys;d=json.loads(sys.stdin.read());print(round(d['value'],1),round(d['roofline']['frac'],4),d['clocks']['sm_mhz'])"
done
