#!/bin/bash
# usage: tools/geom_sweep.sh <workload> "T CS" "T CS" ...   (on the GPU box)
export PYTHONPATH=$PWD
wl=$1; shift
for cfg in "$@"; do set -- $cfg
  echo -n "$wl T=$1 CS=$2: "
  WF_TILE_T=$1 WF_TILE_CS=$2 timeout 300 python bench.py --workload $wl --steps 320 --warmup 32 2>&1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%.4g env-steps/s  %.2f us/step  per-step-launch %.2f us  e2e %.4g' % (d['value'], 1e3*d['ms_per_step'], d['per_step_launch']['us_per_step'], d['e2e']['value']))
except Exception as ex: print('FAILED', ex)"
done
