#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_factored.py -x -q > gpurun_out/r2g_fac.log 2>&1; echo "factored tests rc=$?"
tail -15 gpurun_out/r2g_fac.log
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_factored.py > gpurun_out/r2g_tests.log 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/r2g_tests.log
timeout 600 python bench.py --no-workloads > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; echo "bench rc=$?"
cat gpurun_out/r2g_bench.json; tail -3 gpurun_out/r2g_bench.err
