#!/bin/bash
run() { name=$1; lib=$2; shift 2
  env "$@" MMD_B200_LIB=$lib TAG=$name python tools/kbench.py > gpurun_out/e11_$name.json 2> gpurun_out/e11_$name.err
  cat gpurun_out/e11_$name.json; }
for v in p2 p2a p2b p2c; do
run ${v}_occ1 build_variants/libmmd_$v.so NCH=1184
run ${v} build_variants/libmmd_$v.so NCH=16384
done
