#!/bin/bash
# full GPU suite + bench, both arms
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_tests.log 2>&1; echo "tests rc=$?" 
tail -5 gpurun_out/r2f_tests.log
timeout 600 python bench.py > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench rc=$?"
cat gpurun_out/r2f_bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2f_ref.json 2> gpurun_out/r2f_ref.err; echo "ref rc=$?"
cat gpurun_out/r2f_ref.json
