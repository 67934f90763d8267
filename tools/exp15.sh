#!/bin/bash
# ncu full-set capture of one k_leapfrog launch with one CTA per SM (latency profile) for a variant library
tag=$1; n=${2:-1184}
MMD_B200_LIB=build_variants/libmmd_$tag.so NCH=$n BURN=6 timeout 200 python tools/kbench.py > gpurun_out/e15_${tag}_plain.json 2>&1; cat gpurun_out/e15_${tag}_plain.json
MMD_B200_LIB=build_variants/libmmd_$tag.so NCH=$n BURN=6 timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_leapfrog -s 7 -c 1 -f -o gpurun_out/e15_${tag}_$n python tools/kbench.py > gpurun_out/e15_${tag}_ncu.log 2>&1; tail -n 3 gpurun_out/e15_${tag}_ncu.log
