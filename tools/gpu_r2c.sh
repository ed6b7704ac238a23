#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; cat gpurun_out/r2c_bench.json; tail -5 gpurun_out/r2c_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2c_bench_ref.json 2>> gpurun_out/r2c_bench.err; cat gpurun_out/r2c_bench_ref.json
