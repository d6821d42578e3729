#!/bin/bash
# GPU call 23 of round 2: Float32 resamples through a field on the output grid take the Demons warp kernel -- full GPU suite + A/B
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 900 -x > gpurun_out/r02w_pytest_gpu.log 2>&1
tail -4 gpurun_out/r02w_pytest_gpu.log
for v in off on off on; do
  if [ $v = off ]; then export PLATIPY_B200_WARP_RESAMPLE=0; else unset PLATIPY_B200_WARP_RESAMPLE; fi
  echo "warp_resample_$v $(timeout 200 python profiles/exp_registration_total.py 2>&1 | grep TOTAL)" | tee -a gpurun_out/r02w_ab_warp_resample.log
done
