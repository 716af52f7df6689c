#!/bin/bash
# Round-2 ncu evidence for the registration tail after the late kernel changes (cluster Sinkhorn, dual-form point
# Sinkhorn, TMA attention, patch scores, fine correspondences).  Same conventions as capture_profiles.sh.
set -x
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active
ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r2b_launches_pairs_step.csv python bench.py --workload pairs --pairs 16 --ncu > gpurun_out/r2b_ncu_pairs.log 2>&1
F="--set full --clock-control none --import-source on"
ncu $F -k regex:attention_tma -c 2 -s 16 -o gpurun_out/r2b_full_attention_tma python bench.py --workload pairs --pairs 16 --ncu > /dev/null 2>&1
ncu $F -k regex:"patch_scores|fine_corr_find" -c 2 -s 2 -o gpurun_out/r2b_full_patch_fine python bench.py --workload pairs --pairs 16 --ncu > /dev/null 2>&1
ncu $F -k regex:query_self -c 2 -s 1 -o gpurun_out/r2b_full_radius_self python bench.py --workload descriptor --ncu > /dev/null 2>&1
ls -la gpurun_out/r2b_*
