#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 900 -x > gpurun_out/r02o_pytest_gpu.log 2>&1
tail -3 gpurun_out/r02o_pytest_gpu.log
for v in copy_off copy_on copy_off copy_on; do
  if [ $v = copy_off ]; then export PLATIPY_B200_IDENTITY_COPY=0; else unset PLATIPY_B200_IDENTITY_COPY; fi
  echo "$v $(timeout 200 python profiles/exp_registration_total.py 2>&1 | grep TOTAL)" | tee -a gpurun_out/r02o_ab_identity_copy.log
done
