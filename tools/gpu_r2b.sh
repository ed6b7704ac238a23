#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_gputests.log 2>&1; tail -15 gpurun_out/r2b_gputests.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; cat gpurun_out/r2b_bench.json; tail -3 gpurun_out/r2b_bench.err
