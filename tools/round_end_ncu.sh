#!/bin/bash
# the ncu part of tools/round_end.sh alone: round_end_ncu.sh gri | nc7   (one report per call: a call may
# bring back at most 64 MiB)
mkdir -p gpurun_out
if [ "$1" = gri ]; then
ncu --set full --import-source on --clock-control none --kernel-name regex:k_eval --launch-skip 1 --launch-count 1 -f \
    -o gpurun_out/prof_final python tools/sweep.py --n 65536 --configs 8:384:0 --reps 1 > gpurun_out/ncu_final.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
else
ncu --set full --clock-control none --kernel-name regex:k_eval --launch-skip 1 --launch-count 1 -f \
    -o gpurun_out/prof_nc7 python tools/mech_sweep.py --cases nc7:9472 --reps 1 > gpurun_out/ncu_nc7.log 2>&1
fi
ls -la gpurun_out/
