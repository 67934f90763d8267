#!/bin/bash
# experiment batch 2: L1 prefetch in the linearisation sweeps; stall profile at 1 chain per tile (I-cache hypothesis)
run() { name=$1; lib=$2; shift 2
  env "$@" MMD_B200_LIB=$lib NCH=16384 TAG=$name python tools/kbench.py > gpurun_out/e2_$name.json 2> gpurun_out/e2_$name.err
  cat gpurun_out/e2_$name.json; }
M=manifold_mcmc_for_diffusions_b200/libmmd_b200.so
run main $M
for v in 1 2 4; do run pl1_$v build_variants/libmmd_pl1_$v.so; done
MET=smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__icc_request_hit_rate.pct,gcc__average_cache_request_hit_rate.pct,smsp__inst_executed.sum,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct
for c in 1 8; do
  MMD_CPB=$c NCH=8192 BURN=6 ncu --metrics $MET --clock-control none -k regex:k_leapfrog -s 7 -c 1 --csv --log-file gpurun_out/e2_cpb${c}_ncu.csv python tools/kbench.py > /dev/null 2>&1
  grep -v "^==" gpurun_out/e2_cpb${c}_ncu.csv | awk -F'","' 'NR>1 {gsub(/"/,"",$NF); print $(NF-2), $NF}'
  echo ---
done
