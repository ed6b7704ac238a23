#!/bin/bash
# A/B timing of development builds (tools/devbuild.sh): tools/ab.sh NAME... -> states/s of eval_jacob per build
for name in "$@"; do
  echo "== $name"
  PYJAC_B200_LIB=pyjac_b200/_build/dev_$name.so timeout 120 python tools/sweep.py --n 262144 --configs ${CFG:-8:384:0} --reps 7 2>&1 | grep -E "gs=|error|Error"
done
