#!/bin/bash
# streams / pairs sweep of bench.py (prints value, ms_per_step, e2e)
for cfg in ${SWEEP:-"64 2" "64 4" "64 6" "64 8" "128 8" "256 8"}; do
  set -- $cfg
  timeout 300 python bench.py --pairs $1 --streams $2 --steps 6 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('pairs $1 streams $2', round(d['value']), round(d['ms_per_step'],2), round(d['e2e']['value']))"
done
