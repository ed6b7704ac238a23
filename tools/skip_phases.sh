# per-phase wall time of k_eval by skipping phases (results are wrong on purpose; timing only).  Needs a -DPJ_DEV build:
#   tools/devbuild.sh dev -DPJ_DEV; tools/skip_phases.sh
# masks: 1 A1, 2 B, 4 C, 8 class S, 16 class D, 32 class T (never 64 alone: class T waits for warp 0's arrival), 128 no stores
for m in 0 1 2 4 8 16 32 56 128 136 255; do echo "skip=$m"; PYJAC_B200_LIB=pyjac_b200/_build/dev_dev.so PYJAC_DEBUG_SKIP=$m timeout 100 python tools/sweep.py --configs 8:384:0 --n 131072 --reps 3 2>&1 | grep "gs=8"; done
