#!/bin/bash
run() { name=$1; lib=$2; shift 2
  env "$@" MMD_B200_LIB=$lib TAG=$name python tools/kbench.py > gpurun_out/e9_$name.json 2> gpurun_out/e9_$name.err
  cat gpurun_out/e9_$name.json; }
for n in 1184 16384; do
NCH=$n MMD_B200_LIB=build_variants/libmmd_phase2.so python tools/phase_times.py > gpurun_out/e9_phase_$n.json 2> gpurun_out/e9_phase_$n.err; python -c "
import json; j=json.load(open('gpurun_out/e9_phase_$n.json')); print($n, j['chain_steps_per_s'], j['step_cycles'], j['solver_iterations_per_cta_step']); print(j['cycles_per_cta_step']); print(j['solver_iteration_detail_cycles_per_iteration'])"
done
run pf4_occ1 build_variants/libmmd_pf4.so NCH=1184
run pf4 build_variants/libmmd_pf4.so NCH=16384
run pf4nl2_occ1 build_variants/libmmd_pf4nl2.so NCH=1184
run pf4nl2 build_variants/libmmd_pf4nl2.so NCH=16384
