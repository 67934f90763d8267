#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_notebook_known_answer.py tests/test_gpu_mici_surface.py tests/test_gpu_nuts.py -x -q 2>&1 | tail -n 2
# ncu full-set capture of the FIRST TIMED k_leapfrog launch of tools/kbench.py (BURN = 6 burn-in launches are skipped:
# step size 0.1, chains in the typical set), 16,384 chains
NCH=16384 BURN=6 timeout 200 python tools/kbench.py > gpurun_out/r2g_plain.json 2>&1; cat gpurun_out/r2g_plain.json
NCH=16384 BURN=6 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_leapfrog -s 6 -c 1 -f -o gpurun_out/r2g_full python tools/kbench.py > gpurun_out/r2g_full.log 2>&1; tail -n 2 gpurun_out/r2g_full.log
