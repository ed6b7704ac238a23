#!/bin/bash
# development: eval_jacob throughput for several settings of the phase-DE cost model (plan.py)
# usage: cost_sweep.sh DOTS "S_STEP,S_OVF,D_ITEM,D_COL,T_ITEM,T_IT" ...
dots=$1; shift
for c in "$@"; do echo "dots=$dots costs=$c"; PYJAC_COST_DOTS=$dots PYJAC_COSTS=$c python tools/sweep.py --n 262144 --configs 8:384:0 --reps 5 2>&1 | grep gs=; done
