#!/bin/bash
run() { name=$1; lib=$2; shift 2
  env "$@" MMD_B200_LIB=$lib NCH=16384 TAG=$name python tools/kbench.py > gpurun_out/e5_$name.json 2> gpurun_out/e5_$name.err
  cat gpurun_out/e5_$name.json; }
M=manifold_mcmc_for_diffusions_b200/libmmd_b200.so
run main $M
run mr104 build_variants/libmmd_mr104.so
run mr112 build_variants/libmmd_mr112.so
nproc; numactl -H 2>/dev/null | head -5; nvidia-smi topo -m 2>/dev/null | head -8
python bench.py --e2e-pipeline 4 --no-cpu-baseline > gpurun_out/e5_bench.json 2> gpurun_out/e5_bench.err; python -c "
import json
d=json.loads(open('gpurun_out/e5_bench.json').read().strip().splitlines()[-1])
print(round(d['value']), d['e2e']['value'], d['config'].get('cpu_affinity'), d['clocks'])"
