for m in 0 1 2 4 8 16 32 56; do echo "skip=$m"; PYJAC_DEBUG_SKIP=$m timeout 200 python tools/mech_sweep.py --cases nc7:18944 --reps 2 2>&1 | grep "nc7"; done
