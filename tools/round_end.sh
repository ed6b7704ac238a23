#!/bin/bash
# One GPU-box pass that produces what profiles/ is written from: tests, throughput over mechanism
# sizes, the bench lines.  The ncu captures are made by tools/round_end_ncu.sh (one report per call).
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/gputests_final.log 2>&1; tail -3 gpurun_out/gputests_final.log
python tools/sweep.py --n 262144 --configs 8:384:0,8:512:0 --reps 7 > gpurun_out/sweep_final.log 2>&1; cat gpurun_out/sweep_final.log
python tools/mech_sweep.py --cases gri30:262144,usc2:75776,nc7:18944 > gpurun_out/mech_sweep.md 2> gpurun_out/mech_sweep.err; cat gpurun_out/mech_sweep.md; tail -2 gpurun_out/mech_sweep.err
python tools/batch_sweep.py > gpurun_out/batch_sweep.md 2>&1; cat gpurun_out/batch_sweep.md
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_final.json 2> gpurun_out/bench_final.err
python bench.py > gpurun_out/bench_final.json 2>> gpurun_out/bench_final.err; cat gpurun_out/bench_final.json
for w in usc2 nc7; do python bench.py --workload $w --steps 5 --no-cpu > gpurun_out/bench_$w.json 2>> gpurun_out/bench_final.err; cat gpurun_out/bench_$w.json; done
