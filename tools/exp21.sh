#!/bin/bash
run() { name=$1; lib=$2; shift 2
  env "$@" MMD_B200_LIB=$lib TAG=$name timeout 150 python tools/kbench.py > gpurun_out/e21_$name.json 2> gpurun_out/e21_$name.err
  cat gpurun_out/e21_$name.json; }
L=build_variants/libmmd_carve.so
for r in 1 2; do
run default_$r $L NCH=16384
run c64_$r $L NCH=16384 MMD_SMEM_CARVEOUT=64
run c72_$r $L NCH=16384 MMD_SMEM_CARVEOUT=72
run c86_$r $L NCH=16384 MMD_SMEM_CARVEOUT=86
run c100_$r $L NCH=16384 MMD_SMEM_CARVEOUT=100
done
