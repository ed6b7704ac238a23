#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2h_tests.log 2>&1; echo "tests rc=$?"
tail -6 gpurun_out/r2h_tests.log
timeout 600 python tools/measure_gates.py > gpurun_out/r2h_gates.md 2>&1; cat gpurun_out/r2h_gates.md
timeout 600 python tools/batch_tlb.py > gpurun_out/r2h_stride.md 2>&1; cat gpurun_out/r2h_stride.md
timeout 600 python bench.py --no-workloads --no-cpu > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r2h_bench.json')); print(json.dumps({k:d[k] for k in ('value','e2e','e2e_factored','consumer')}))"
