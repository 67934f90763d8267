#!/bin/bash
L=build_variants/libmmd_$1.so
MMD_B200_LIB=$L timeout 150 python -m pytest tests/test_gpu_parity_small.py -x -v 2>&1 | tail -n 25 | cut -c1-200
echo "rc=$?"
