#!/bin/bash
# usage: scratch/ab_bench.sh VAR v1 v2 ...  -> runs bench.py (no CPU baseline) once per value of the environment variable
VAR=$1; shift
for v in "$@"; do
  env $VAR=$v timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > /tmp/ab.json
  python -c "
import json; d=json.load(open('/tmp/ab.json')); print('$VAR=$v', round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1))"
done
