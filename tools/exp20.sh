#!/bin/bash
echo skip-tests
run() { name=$1; lib=$2; shift 2
  env "$@" MMD_B200_LIB=$lib TAG=$name timeout 150 python tools/kbench.py > gpurun_out/e20_$name.json 2> gpurun_out/e20_$name.err
  cat gpurun_out/e20_$name.json; }
M=manifold_mcmc_for_diffusions_b200/libmmd_b200.so
for r in 1 2 3; do
run main_$r $M NCH=16384
run $1_$r build_variants/libmmd_$1.so NCH=16384
done
