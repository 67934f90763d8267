#!/bin/bash
MMD_B200_LIB=build_variants/libmmd_$1.so timeout 600 python -m pytest tests/test_gpu_parity_small.py tests/test_gpu_golden_canonical.py tests/test_gpu_parity_variants.py tests/test_gpu_bundled_configs.py tests/test_gpu_sir.py -x -q 2>&1 | tail -n 3
run() { name=$1; lib=$2; shift 2
  env "$@" MMD_B200_LIB=$lib TAG=$name timeout 150 python tools/kbench.py > gpurun_out/e22_$name.json 2> gpurun_out/e22_$name.err
  cat gpurun_out/e22_$name.json; }
M=manifold_mcmc_for_diffusions_b200/libmmd_b200.so
for r in 1 2 3; do
run main_$r $M NCH=16384
run $1_$r build_variants/libmmd_$1.so NCH=16384
done
