#!/bin/bash
# round 2: the ncu evidence profiles/r02_* is written from (one GPU).  usage: round2_end.sh
mkdir -p gpurun_out
M=gpu__time_duration.sum,smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,dram__bytes_write.sum,dram__bytes_read.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__inst_executed_op_shared_ld.sum,smsp__inst_executed_op_shared_st.sum,smsp__inst_executed_op_global_ld.sum,smsp__inst_executed_op_global_st.sum,launch__registers_per_thread,launch__occupancy_limit_registers,sm__warps_active.avg.pct_of_peak_sustained_active
timeout 600 ncu --set full --import-source on --clock-control none --kernel-name regex:k_eval --launch-skip 1 --launch-count 1 -f \
    -o gpurun_out/prof_r02_jac python tools/sweep.py --n 65536 --configs 8:384:0 --reps 1 > gpurun_out/r02_ncu_jac.log 2>&1; tail -2 gpurun_out/r02_ncu_jac.log
timeout 600 ncu --metrics $M --clock-control none --kernel-name regex:"k_eval|k_jvp|k_newton" --launch-skip 3 --launch-count 3 \
    python tools/fac_once.py 65536 > gpurun_out/r02_ncu_consumers.log 2>&1; grep -E "k_eval|k_jvp|k_newton|duration|inst_executed|wavefronts|issue_active|dram__|fp64|registers|warps_active" gpurun_out/r02_ncu_consumers.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 2 --warmup 1 > gpurun_out/r02_bench_under_ncu.log 2>&1; tail -c 600 gpurun_out/r02_bench_under_ncu.log; wc -l gpurun_out/r02_launches.csv
timeout 600 python tools/batch_tlb.py > gpurun_out/r02_stride.md 2>&1; cat gpurun_out/r02_stride.md
ls -la gpurun_out/prof_r02_jac.ncu-rep
