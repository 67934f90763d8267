#!/bin/bash
# Final round-2 artefacts on one B200 for the built library: GPU tests, bench line, kernel timing, ncu full-set
# capture of one k_leapfrog launch, ncu launch list of the bench command.
tag=${1:-r2f}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_gputest.log 2>&1; echo "gputest rc=$?"; tail -n 2 gpurun_out/${tag}_gputest.log
timeout 600 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/${tag}_bench.json
NCH=16384 TAG=$tag timeout 200 python tools/kbench.py > gpurun_out/${tag}_kbench.json 2> gpurun_out/${tag}_kbench.err; cat gpurun_out/${tag}_kbench.json
bash tools/op_times_grid.sh
NCH=16384 BURN=6 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_leapfrog -s 5 -c 1 -f -o gpurun_out/${tag}_full python tools/kbench.py > gpurun_out/${tag}_full.log 2>&1; tail -n 2 gpurun_out/${tag}_full.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --steps 8 --warmup 8 --burnin 4 --no-cpu-baseline --e2e-steps 1 > gpurun_out/${tag}_ncu_bench.log 2>&1; echo "ncu rc=$?"
