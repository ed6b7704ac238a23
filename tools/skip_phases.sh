# per-phase wall time of k_jacobian by skipping phases (results are wrong on purpose; timing only)
for m in 0 1 2 4 8 16 32 56; do echo "skip=$m"; PYJAC_DEBUG_SKIP=$m timeout 100 python tools/sweep.py --configs 8:512:0 --n 131072 --reps 3 2>&1 | grep "gs=8"; done
