#!/bin/bash
# A/B of development builds over mechanism sizes: tools/ab_mech.sh CASES NAME...  ("main" = the in-tree library)
cases=$1; shift
for name in "$@"; do
  echo "== $name"
  if [ "$name" = main ]; then unset PYJAC_B200_LIB; else export PYJAC_B200_LIB=pyjac_b200/_build/dev_$name.so; fi
  timeout 300 python tools/mech_sweep.py --cases $cases --reps 3 2>&1 | grep -E "^\| (gri30|usc2|nc7)|rror"
done
