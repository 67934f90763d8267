#!/bin/bash
# experiment batch 1: first-pass K skip, 4 CTAs/SM, persisting-L2 carve-out
run() { # name lib env...
  name=$1; lib=$2; shift 2
  env "$@" MMD_B200_LIB=$lib NCH=16384 TAG=$name python tools/kbench.py > gpurun_out/e1_$name.json 2> gpurun_out/e1_$name.err
  cat gpurun_out/e1_$name.json
}
M=manifold_mcmc_for_diffusions_b200/libmmd_b200.so
run main $M
run nofps build_variants/libmmd_nofps.so
run minb4 build_variants/libmmd_minb4.so
run main_l2p32 $M MMD_L2_PERSIST_MB=32
run main_l2p64 $M MMD_L2_PERSIST_MB=64
run main_l2p96 $M MMD_L2_PERSIST_MB=96
run main2 $M
python -m pytest tests/test_gpu_golden_canonical.py tests/test_gpu_parity_small.py tests/test_gpu_newton.py tests/test_gpu_sir.py -x -q -m gpu 2>&1 | tail -3
