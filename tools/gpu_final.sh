#!/bin/bash
# round-end pass on one GPU: the suite, both bench arms, the launch list of the bench command
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_tests_final.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r02_tests_final.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err; echo "ref rc=$?"; cat gpurun_out/r02_bench_ref.json
timeout 900 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; echo "bench rc=$?"; cat gpurun_out/r02_bench.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 2 --warmup 1 > gpurun_out/r02_bench_under_ncu.log 2>&1; wc -l gpurun_out/r02_launches.csv
python -c "import __graft_entry__ as g; g.smoke()"
