#!/bin/bash
for n in 1184 16384; do
NCH=$n MMD_B200_LIB=build_variants/libmmd_fp2phase.so python tools/phase_times.py > gpurun_out/e8_phase_$n.json 2> gpurun_out/e8_phase_$n.err; python -c "
import json; j=json.load(open('gpurun_out/e8_phase_$n.json')); print($n, j['chain_steps_per_s'], j['step_cycles'], j['solver_iterations_per_cta_step']); print(j['cycles_per_cta_step'])"
done
