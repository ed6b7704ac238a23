#!/bin/bash
# round 2, first GPU pass of the record-stream kernel: parity, A/B against the schedule-table kernel, counters
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "stream_kernel or two_handles" > gpurun_out/r2a_streams.log 2>&1; tail -15 gpurun_out/r2a_streams.log
timeout 300 python tools/sweep.py --n 262144 --configs 8:384:0:1,8:384:0:0,8:256:0:1,4:384:0:1 --reps 7 > gpurun_out/r2a_sweep.log 2>&1; cat gpurun_out/r2a_sweep.log
M=gpu__time_duration.sum,smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,dram__bytes_write.sum,dram__bytes_read.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,lts__t_sectors_srcunit_tex_op_read.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum
timeout 300 ncu --metrics $M --clock-control none -k regex:k_jac6 -s 1 -c 1 python tools/sweep.py --n 65536 --configs 8:384:0:1 --reps 1 > gpurun_out/r2a_ncu_metrics.log 2>&1; grep -E "k_jac6|duration|inst_executed|wavefronts|issue_active|bank_conflicts|dram__|fp64|lts__|requests" gpurun_out/r2a_ncu_metrics.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_jac6 -s 1 -c 1 -f -o gpurun_out/prof_r2a python tools/sweep.py --n 65536 --configs 8:384:0:1 --reps 1 > gpurun_out/r2a_ncu_full.log 2>&1; tail -3 gpurun_out/r2a_ncu_full.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_gputests.log 2>&1; tail -8 gpurun_out/r2a_gputests.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; cat gpurun_out/r2a_bench.json; tail -3 gpurun_out/r2a_bench.err
ls -la gpurun_out | head -30
