#!/bin/bash
# GPU call 26 of round 2: the force kernel's last block closes the iteration (no finish launch) -- GPU suite, A/B, sanitizer over the hot path
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 900 -x > gpurun_out/r02z_pytest_gpu.log 2>&1
tail -4 gpurun_out/r02z_pytest_gpu.log
for v in off on off on; do
  if [ $v = off ]; then export PLATIPY_B200_FINISH_IN_FORCE=0; else unset PLATIPY_B200_FINISH_IN_FORCE; fi
  echo "finish_in_force_$v $(timeout 200 python profiles/exp_registration_total.py 2>&1 | grep TOTAL)" | tee -a gpurun_out/r02z_ab_finish_in_force.log
done
unset PLATIPY_B200_FINISH_IN_FORCE
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck synccheck; do
  timeout 600 $CS --tool $tool --print-limit 20 python profiles/sanitize_hot_path.py > gpurun_out/r02z_sanitizer_${tool}_hot_path.log 2>&1
  echo "$tool hot path: $(grep -c 'SANITIZE RUN OK' gpurun_out/r02z_sanitizer_${tool}_hot_path.log) ok; $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/r02z_sanitizer_${tool}_hot_path.log | tail -1)"
done
