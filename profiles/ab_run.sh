#!/bin/bash
# usage (on the GPU box): bash profiles/ab_run.sh [extra bench.py flags]   -> alternating HEAD / working-tree bench runs
cd "$(dirname "$0")/.."
for i in 1 2; do
  for v in head new; do
    if [ $v = head ]; then export HPFG_B200_LIB=$PWD/profiles/ab/libhead.so; else unset HPFG_B200_LIB; fi
    timeout 180 python bench.py --steps 20 --warmup 5 --no-cpu "$@" 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', round(d['ms_per_step'],4), round(d['e2e']['ms_per_step'],4))"
  done
done
