#!/bin/bash
MMD_B200_LIB=build_variants/libmmd_$1.so timeout 500 python -m pytest tests/test_gpu_parity_small.py tests/test_gpu_golden_canonical.py tests/test_gpu_parity_variants.py tests/test_gpu_newton.py tests/test_gpu_bundled_configs.py -x -q 2>&1 | tail -n 3
shift
bash tools/exp16.sh "$@"
