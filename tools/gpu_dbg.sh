#!/bin/bash
export PYJAC_B200_LIB=pyjac_b200/_build/dev_rel4.so
timeout 300 python tools/sweep.py --n 262144 --configs 8:384:0:0,8:384:0:1 --reps 7 2>&1 | grep -E "gs=|rror"
M=gpu__time_duration.sum,smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed
timeout 300 ncu --metrics $M --clock-control none -k regex:k_eval -s 1 -c 1 python tools/sweep.py --n 65536 --configs 8:384:0:0 --reps 1 2>&1 | grep -E "duration|inst_executed|wavefronts|bank_conflicts"
