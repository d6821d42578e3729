#!/bin/bash
# GPU call 31 of round 2: running output pointers in the fused smoothing kernel (13 % fewer instructions) -- A/B against the previous build,
# with the plain pass compiled for 5 CTAs / SM (72 registers, 16 B spilled) and for 4 (86 registers); then the smoothing / Demons parity tests
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python profiles/ab_variants.py prev=lib=libb200reg_prevptr.so ptr5=lib=libb200reg.so ptr4=lib=libb200reg_ctas4.so prev2=lib=libb200reg_prevptr.so ptr5b=lib=libb200reg.so ptr4b=lib=libb200reg_ctas4.so 2>&1 | grep -v "^AB " | tee gpurun_out/r02ac_ab_running_pointers.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -p no:cacheprovider --timeout 900 -x > gpurun_out/r02ac_pytest_gpu.log 2>&1
tail -3 gpurun_out/r02ac_pytest_gpu.log
