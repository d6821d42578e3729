#!/bin/bash
# GPU call 34 of round 2: the committed head once more -- GPU suite, smoke(), memcheck over the hot-path script (now with a single-label call)
# and over the test of the single-image resampling kernels
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 900 > gpurun_out/r02_final3_pytest_gpu.log 2>&1
tail -3 gpurun_out/r02_final3_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -2 | tee gpurun_out/r02_final3_smoke.log
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 600 $CS --tool memcheck --print-limit 20 python profiles/sanitize_hot_path.py > gpurun_out/r02_final3_sanitizer_memcheck_hot_path.log 2>&1
echo "memcheck hot path: $(grep -c 'SANITIZE RUN OK' gpurun_out/r02_final3_sanitizer_memcheck_hot_path.log) ok; $(grep -E 'ERROR SUMMARY' gpurun_out/r02_final3_sanitizer_memcheck_hot_path.log | tail -1)"
timeout 600 $CS --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "float32_resample_through" > gpurun_out/r02_final3_sanitizer_memcheck_single_image_kernels.log 2>&1
echo "memcheck single-image kernels: $(grep -E 'passed|failed' gpurun_out/r02_final3_sanitizer_memcheck_single_image_kernels.log | tail -1); $(grep -E 'ERROR SUMMARY' gpurun_out/r02_final3_sanitizer_memcheck_single_image_kernels.log | tail -1)"
