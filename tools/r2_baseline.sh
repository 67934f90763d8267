#!/bin/bash
# Round-2 baseline artefacts on one B200: GPU tests, bench line, ncu launch list of the same command, stream
# microbenchmark, kernel-level timing.  Outputs under gpurun_out/ (copied into profiles/ by hand).
tag=${1:-r2d}
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_gputest.log 2>&1; echo "gputest rc=$?"
tail -2 gpurun_out/${tag}_gputest.log
python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
cut -c1-300 gpurun_out/${tag}_bench.json
NCH=16384 TAG=$tag python tools/kbench.py > gpurun_out/${tag}_kbench.json 2> gpurun_out/${tag}_kbench.err; cat gpurun_out/${tag}_kbench.json
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/stream_bench tools/stream_bench.cu && /tmp/stream_bench > gpurun_out/${tag}_stream_bench.txt 2>&1; cat gpurun_out/${tag}_stream_bench.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --steps 8 --warmup 8 --burnin 4 --no-cpu-baseline --e2e-steps 1 > gpurun_out/${tag}_ncu_bench.log 2>&1; echo "ncu rc=$?"
