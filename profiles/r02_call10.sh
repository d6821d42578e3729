#!/bin/bash
# GPU call 10 of round 2 (1 GPU): running output offsets in the smoothing kernels (A/B against the previous build), suite.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 400 python profiles/ab_variants.py prev=lib=libb200reg_prev.so new= prev_again=lib=libb200reg_prev.so new_again= > gpurun_out/r02j_ab_running_offsets.log 2>&1
grep -v "^AB" gpurun_out/r02j_ab_running_offsets.log | cut -c1-250
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 -x > gpurun_out/r02j_pytest_gpu.log 2>&1
tail -3 gpurun_out/r02j_pytest_gpu.log
