#!/bin/bash
# GPU call 8 of round 2 (1 GPU): fast mode against parity mode (small size first, then the headline size), suite.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 200 python profiles/exp_fast_mode.py 128 96 64 > gpurun_out/r02h_fast_mode_small.log 2>&1; tail -c 1500 gpurun_out/r02h_fast_mode_small.log
timeout 300 python profiles/exp_fast_mode.py > gpurun_out/r02h_fast_mode.log 2>&1; tail -c 1800 gpurun_out/r02h_fast_mode.log
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 -x > gpurun_out/r02h_pytest_gpu.log 2>&1
tail -4 gpurun_out/r02h_pytest_gpu.log
