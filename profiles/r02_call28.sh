#!/bin/bash
# GPU call 28 of round 2: compile-time radii in the Float32 Gaussian passes -- GPU suite + A/B
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 900 -x > gpurun_out/r02ab_pytest_gpu.log 2>&1
tail -4 gpurun_out/r02ab_pytest_gpu.log
for v in runtime static runtime static; do
  if [ $v = runtime ]; then export PLATIPY_B200_CONV_STATIC_RADIUS=0; else unset PLATIPY_B200_CONV_STATIC_RADIUS; fi
  echo "radius_$v $(timeout 200 python profiles/exp_registration_total.py 2>&1 | grep TOTAL)" | tee -a gpurun_out/r02ab_ab_static_radius.log
done
