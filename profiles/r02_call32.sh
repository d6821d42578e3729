#!/bin/bash
# GPU call 32 of round 2: the GPU suite and smoke() on the committed build; memcheck over the tests of the kernels added last
# (restricted pyramid levels at odd shapes, the warp-kernel resample)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 900 > gpurun_out/r02_final2_pytest_gpu.log 2>&1
tail -3 gpurun_out/r02_final2_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -2 | tee gpurun_out/r02_final2_smoke.log
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 900 $CS --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "restricted or smooth_and_resample or float32_resample_through" > gpurun_out/r02_final2_sanitizer_memcheck_new_kernels.log 2>&1
echo "memcheck new kernels: $(grep -E 'passed|failed' gpurun_out/r02_final2_sanitizer_memcheck_new_kernels.log | tail -1); $(grep -E 'ERROR SUMMARY' gpurun_out/r02_final2_sanitizer_memcheck_new_kernels.log | tail -1)"
