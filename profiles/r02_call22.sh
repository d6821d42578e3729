#!/bin/bash
# GPU call 22 of round 2: restricted pyramid levels with the adjacent-pair fast path -- parity subset + A/B over the pay-off threshold
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider --timeout 600 -x -k "restricted or smooth_and_resample or fast_symmetric" > gpurun_out/r02v_pytest_restricted.log 2>&1
tail -4 gpurun_out/r02v_pytest_restricted.log
for v in 0 0.6 1.6 0 0.6 1.6; do
  export PLATIPY_B200_PYRAMID_RESTRICT_COST=$v
  echo "cost<=$v $(timeout 200 python profiles/exp_registration_total.py 2>&1 | grep TOTAL)" | tee -a gpurun_out/r02v_ab_pyramid_restrict.log
done
