#!/bin/bash
# compute-sanitizer over the hot path at small sizes: memcheck, racecheck (shared-memory hazards: the fused smoothing
# kernel's staged tiles, the Deriche tiles, the block-private STAPLE histogram), synccheck (barriers / mbarrier use); then memcheck and
# racecheck over the morphology (union-find CCL) and Mattes (histogram atomics) GPU tests.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck synccheck; do
  timeout 600 $CS --tool $tool --print-limit 20 python profiles/sanitize_hot_path.py > gpurun_out/r02g_sanitizer_${tool}_hot_path.log 2>&1
  echo "$tool hot path: $(grep -c 'SANITIZE RUN OK' gpurun_out/r02g_sanitizer_${tool}_hot_path.log) ok; $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/r02g_sanitizer_${tool}_hot_path.log | tail -1)"
done
for tool in memcheck racecheck; do
  timeout 900 $CS --tool $tool --print-limit 20 python -m pytest tests/test_gpu_morph.py tests/test_gpu_linear.py -q -p no:cacheprovider -k "not fullsize" > gpurun_out/r02g_sanitizer_${tool}_morph_linear.log 2>&1
  echo "$tool morph+linear: $(grep -E 'passed|failed' gpurun_out/r02g_sanitizer_${tool}_morph_linear.log | tail -1); $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/r02g_sanitizer_${tool}_morph_linear.log | tail -1)"
done
