#!/bin/bash
# Build a variant of libmmd_b200.so with extra -D flags for the FHN translation unit and the API
# (mmd_ops_fhn_r5.cu, the instantiation the bench workload runs; the other model units are taken from build/): tools/build_variant.sh NAME -DMMD_PHASE_CLOCK ...
set -e
cd "$(dirname "$0")/.."
name=$1; shift
out=build_variants/$name
mkdir -p $out
F="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC"
C=manifold_mcmc_for_diffusions_b200/csrc
nvcc $F "$@" -c -o $out/mmd_api.o $C/mmd_api.cu &
nvcc $F "$@" -c -o $out/mmd_ops_fhn_r5.o $C/mmd_ops_fhn_r5.cu &
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build_variants/libmmd_$name.so $out/mmd_api.o $out/mmd_ops_fhn_r5.o \
  build/mmd_ops_fhn.o build/mmd_ops_fhn_r16.o build/mmd_ops_sir.o
ls -la build_variants/libmmd_$name.so
