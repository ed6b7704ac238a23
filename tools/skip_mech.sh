#!/bin/bash
# per-phase wall time by skipping phases (results are wrong on purpose; timing only): skip_mech.sh CASE
# masks: 1 A1, 2 B, 4 C, 8 class S, 16 class D, 32 class T, 56 all of DE
for m in 0 1 2 4 8 16 32 56; do echo "skip=$m"; PYJAC_DEBUG_SKIP=$m timeout 200 python tools/mech_sweep.py --cases $1 --reps 2 2>&1 | grep -E "^\| (gri30|usc2|nc7)"; done
