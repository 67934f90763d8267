#!/bin/bash
run() { name=$1; lib=$2; shift 2
  env "$@" MMD_B200_LIB=$lib NCH=16384 TAG=$name python tools/kbench.py > gpurun_out/e6_$name.json 2> gpurun_out/e6_$name.err
  cat gpurun_out/e6_$name.json; }
M=manifold_mcmc_for_diffusions_b200/libmmd_b200.so
run ca $M
run ca_l2pf0 build_variants/libmmd_ca_l2pf0.so
run ca_nohint build_variants/libmmd_ca_nohint.so
run ca_pf3 build_variants/libmmd_ca_pf3.so
run ca2 $M
