#!/bin/bash
run() { name=$1; lib=$2; shift 2
  env "$@" MMD_B200_LIB=$lib TAG=$name timeout 120 python tools/kbench.py > gpurun_out/e16_$name.json 2> gpurun_out/e16_$name.err
  cat gpurun_out/e16_$name.json; }
for r in 1 2; do
for v in "$@"; do
run ${v}_$r build_variants/libmmd_$v.so NCH=16384
done
done
