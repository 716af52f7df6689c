#!/bin/bash
# Round-2 ncu evidence (run under gpurun on one B200; outputs go to gpurun_out/, summaries are then written into
# profiles/ by scripts/launch_summary.py, scripts/traffic_from_launches.py and profiles/ncu_summary.py).
# Numbers printed by runs under ncu are never bench values.
set -x
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active
# every launch of a descriptor step / a registration step (warm step + measured step; the summaries use the last step)
ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r2_launches_descriptor_step.csv python bench.py --workload descriptor --ncu > gpurun_out/r2_ncu_desc.log 2>&1
ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r2_launches_pairs_step.csv python bench.py --workload pairs --pairs 16 --ncu > gpurun_out/r2_ncu_pairs.log 2>&1
# full captures of the top kernels
F="--set full --clock-control none --import-source on"
ncu $F -k regex:gemm_tf32x3_ts -c 1 -s 2 -o gpurun_out/r2_full_gemm_k1920_n128 python scripts/debug/gemm_ncu.py 130000 1920 128 > /dev/null 2>&1
ncu $F -k regex:gemm_tf32x3_ts -c 1 -s 2 -o gpurun_out/r2_full_gemm_k3840_n256 python scripts/debug/gemm_ncu.py 46000 3840 256 > /dev/null 2>&1
ncu $F -k regex:gemm_tf32x3_ts -c 1 -s 2 -o gpurun_out/r2_full_gemm_k480_n32 python scripts/debug/gemm_ncu.py 909000 480 32 > /dev/null 2>&1
ncu $F -k regex:sinkhorn_patch -c 1 -s 2 -o gpurun_out/r2_full_sinkhorn_patch python scripts/debug/sk_ncu.py > /dev/null 2>&1
ncu $F -k regex:kpconv_gather -c 4 -s 10 -o gpurun_out/r2_full_gather python bench.py --workload descriptor --ncu > /dev/null 2>&1
ncu $F -k regex:query_self -c 2 -s 4 -o gpurun_out/r2_full_radius_self python bench.py --workload descriptor --ncu > /dev/null 2>&1
ncu $F -k regex:attention_tc -c 2 -s 16 -o gpurun_out/r2_full_attention python bench.py --workload pairs --pairs 16 --ncu > /dev/null 2>&1
ls -la gpurun_out/r2_*
