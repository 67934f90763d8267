#!/bin/bash
run() { name=$1; lib=$2; shift 2
  env "$@" MMD_B200_LIB=$lib TAG=$name timeout 150 python tools/kbench.py > gpurun_out/e19_$name.json 2> gpurun_out/e19_$name.err
  cat gpurun_out/e19_$name.json; }
M=manifold_mcmc_for_diffusions_b200/libmmd_b200.so
for r in 1 2; do
run main_$r $M NCH=16384
run keep2_$r build_variants/libmmd_keep2.so NCH=16384
run nt384_$r build_variants/libmmd_nt384.so NCH=16384 MMD_CPB=16
run nt384r88_$r build_variants/libmmd_nt384r88.so NCH=16384 MMD_CPB=16
done
