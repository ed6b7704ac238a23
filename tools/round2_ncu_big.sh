#!/bin/bash
# round 2: ncu --set full of the global-memory plans (USC-II- and n-heptane-shaped mechanisms); only the metric summaries are
# brought back (two reports exceed what one gpurun call may return)
mkdir -p gpurun_out /tmp/rep
for c in usc2:37888 nc7:9472; do
  name=${c%%:*}
  timeout 900 ncu --set full --clock-control none --kernel-name regex:k_eval --launch-skip 1 --launch-count 1 -f \
      -o /tmp/rep/prof_r02_$name python tools/mech_sweep.py --cases $c --reps 1 > gpurun_out/r02_ncu_$name.log 2>&1; tail -1 gpurun_out/r02_ncu_$name.log
  python tools/ncu_summary.py /tmp/rep/prof_r02_$name.ncu-rep > gpurun_out/r02_k_eval_jac_$name.txt
  for k in l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed l1tex__t_sector_hit_rate.pct lts__t_sector_hit_rate.pct; do python tools/ncu_summary.py /tmp/rep/prof_r02_$name.ncu-rep $k | grep -E "^$k " >> gpurun_out/r02_k_eval_jac_$name.txt; done
  sort -u -o gpurun_out/r02_k_eval_jac_$name.txt gpurun_out/r02_k_eval_jac_$name.txt
  cat gpurun_out/r02_k_eval_jac_$name.txt | grep -E "dram__bytes|duration|lsu_wavefronts.avg.pct|hit_rate"
done
