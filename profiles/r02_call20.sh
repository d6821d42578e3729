#!/bin/bash
# GPU call 20 of round 2: restricted pyramid levels (pyramid.cuh) -- parity tests, A/B of the whole registration with the knob off / on
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider --timeout 600 -x -k "restricted or smooth_and_resample" > gpurun_out/r02t_pytest_restricted.log 2>&1
tail -15 gpurun_out/r02t_pytest_restricted.log
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 900 -x > gpurun_out/r02t_pytest_gpu.log 2>&1
tail -3 gpurun_out/r02t_pytest_gpu.log
for v in restrict_off restrict_on restrict_off restrict_on; do
  if [ $v = restrict_off ]; then export PLATIPY_B200_PYRAMID_RESTRICT=0; else unset PLATIPY_B200_PYRAMID_RESTRICT; fi
  echo "$v $(timeout 200 python profiles/exp_registration_total.py 2>&1 | grep TOTAL)" | tee -a gpurun_out/r02t_ab_pyramid_restrict.log
done
