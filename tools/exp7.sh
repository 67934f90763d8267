#!/bin/bash
run() { name=$1; lib=$2; shift 2
  env "$@" MMD_B200_LIB=$lib TAG=$name python tools/kbench.py > gpurun_out/e7_$name.json 2> gpurun_out/e7_$name.err
  cat gpurun_out/e7_$name.json; }
M=manifold_mcmc_for_diffusions_b200/libmmd_b200.so
run main $M NCH=16384
run fp1 build_variants/libmmd_fp1.so NCH=16384
run fp2 build_variants/libmmd_fp2.so NCH=16384
run occ1 $M NCH=1184
run occ2 $M NCH=2368
run occ3 $M NCH=3552
run main2 $M NCH=16384
MMD_B200_LIB=build_variants/libmmd_fp2phase.so python tools/phase_times.py > gpurun_out/e7_fp2_phase.json 2> gpurun_out/e7_fp2_phase.err; cut -c1-1500 gpurun_out/e7_fp2_phase.json
