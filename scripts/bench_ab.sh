#!/bin/bash
# Development aid: bench.py under two settings of one environment knob.   scripts/bench_ab.sh RDFC_UMMA_PDL 1 0
knob=$1; shift
for v in "$@"; do
  env $knob=$v python bench.py 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('$knob=$v', round(d['ms_per_step'], 3), 'ms', round(d['value']), 'maps/s  e2e', round(d['e2e']['value']), ' nlspn frac', round(d['roofline']['frac'], 3), d['clocks'])"
done
