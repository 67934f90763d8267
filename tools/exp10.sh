#!/bin/bash
# usage: exp10.sh TAG  -> parity tests + kbench + phase profile for build_variants/libmmd_TAG.so / libmmd_TAGphase.so
tag=$1
L=build_variants/libmmd_$tag.so
MMD_B200_LIB=$L timeout 400 python -m pytest tests/test_gpu_parity_small.py tests/test_gpu_golden_canonical.py tests/test_gpu_parity_variants.py -x -q 2>&1 | tail -n 3
run() { name=$1; lib=$2; shift 2
  env "$@" MMD_B200_LIB=$lib TAG=$name timeout 120 python tools/kbench.py > gpurun_out/e10_$name.json 2> gpurun_out/e10_$name.err
  cat gpurun_out/e10_$name.json; }
run ${tag}_occ1 $L NCH=1184
run ${tag} $L NCH=16384
run ${tag}_b $L NCH=16384
if [ -f build_variants/libmmd_${tag}phase.so ]; then
for n in 1184 16384; do
NCH=$n MMD_B200_LIB=build_variants/libmmd_${tag}phase.so timeout 120 python tools/phase_times.py > gpurun_out/e10_${tag}_phase_$n.json 2> gpurun_out/e10_${tag}_phase_$n.err; python -c "
import json; j=json.load(open('gpurun_out/e10_${tag}_phase_$n.json')); print($n, j['chain_steps_per_s'], j['step_cycles'], j['solver_iterations_per_cta_step']); print(j['cycles_per_cta_step']); print(json.dumps(j["solver_iteration_detail_cycles_per_iteration"], indent=0))"
done
fi
