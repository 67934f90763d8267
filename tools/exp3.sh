#!/bin/bash
run() { name=$1; lib=$2; shift 2
  env "$@" MMD_B200_LIB=$lib NCH=16384 TAG=$name python tools/kbench.py > gpurun_out/e3_$name.json 2> gpurun_out/e3_$name.err
  cat gpurun_out/e3_$name.json; }
M=manifold_mcmc_for_diffusions_b200/libmmd_b200.so
run inl8 build_variants/libmmd_inl.so
for c in 8 4 2 1; do run noinl$c $M MMD_CPB=$c; done
python -m pytest tests/test_gpu_golden_canonical.py tests/test_gpu_parity_small.py -x -q -m gpu 2>&1 | tail -2
