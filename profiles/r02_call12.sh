#!/bin/bash
# GPU call 12 of round 2 (1 GPU): pyramid-Gaussian kernels with sliding-window inner loops against the previous build (whole registration).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in prev new prev new; do
  if [ $v = prev ]; then export PLATIPY_B200_LIB=$PWD/platipy_b200/libb200reg_prev.so; else unset PLATIPY_B200_LIB; fi
  echo "$v $(timeout 200 python profiles/exp_registration_total.py 2>&1 | grep TOTAL)" | tee -a gpurun_out/r02l_ab_pyramid_kernels.log
done
unset PLATIPY_B200_LIB
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fusion.py tests/test_golden.py -q -p no:cacheprovider -m gpu --timeout 600 -x > gpurun_out/r02l_pytest_gpu.log 2>&1
tail -3 gpurun_out/r02l_pytest_gpu.log
