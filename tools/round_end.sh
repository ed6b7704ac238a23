#!/bin/bash
# One GPU-box pass that produces what profiles/ is written from: tests, throughput over mechanism
# sizes, the ncu --set full capture of the Jacobian kernel, the launch list of bench.py, the bench lines.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/gputests_final.log 2>&1; tail -3 gpurun_out/gputests_final.log
python tools/sweep.py --n 262144 --configs 8:384:0,8:512:0 --reps 7 > gpurun_out/sweep_final.log 2>&1; cat gpurun_out/sweep_final.log
python tools/mech_sweep.py > gpurun_out/mech_sweep.md 2> gpurun_out/mech_sweep.err; cat gpurun_out/mech_sweep.md; tail -2 gpurun_out/mech_sweep.err
ncu --set full --import-source on --clock-control none --kernel-name regex:k_eval --launch-skip 1 --launch-count 1 -f \
    -o gpurun_out/prof_final python tools/sweep.py --n 65536 --configs 8:384:0 --reps 1 > gpurun_out/ncu_final.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_final.json 2> gpurun_out/bench_final.err
python bench.py > gpurun_out/bench_final.json 2>> gpurun_out/bench_final.err; cat gpurun_out/bench_final.json
for w in usc2 nc7; do python bench.py --workload $w --steps 5 --no-cpu > gpurun_out/bench_$w.json 2>> gpurun_out/bench_final.err; cat gpurun_out/bench_$w.json; done
ncu --set full --import-source on --clock-control none --kernel-name regex:k_eval --launch-skip 1 --launch-count 1 -f \
    -o gpurun_out/prof_nc7 python tools/mech_sweep.py --cases nc7:9472 --reps 1 > gpurun_out/ncu_nc7.log 2>&1
