#!/bin/bash
# First GPU call of round 2 (one gpurun invocation, about 6 minutes):
#   1. the GPU tests written after round 1's last GPU minute (tests/test_gpu_zzz_session3.py), then the whole GPU suite;
#   2. A/B of the TMA staging variants of the fused smoothing kernel (row-wise bulk copies, one tensor-map copy per plane tile)
#      against the default cp.async staging -- DVF bit-identity is part of the harness;
#   3. the secondary measurements incl. the rows added in round 1's third session.
# Usage:  gpurun --timeout 900 -- 'bash profiles/r02_first_run.sh'
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_gpu_zzz_session3.py -v -p no:cacheprovider > gpurun_out/r02_pytest_session3.log 2>&1
tail -15 gpurun_out/r02_pytest_session3.log
python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02_pytest_gpu.log 2>&1
tail -3 gpurun_out/r02_pytest_gpu.log
# (prebuild it in the container so that it travels with the snapshot: the same make line)
[ -f platipy_b200/libb200reg_tma.so ] || make -C platipy_b200/csrc OUT=../libb200reg_tma.so EXTRA="-DB200REG_ENABLE_ZM_TMA -DB200REG_AB_VARIANTS" > gpurun_out/r02_build_tma.log 2>&1
python profiles/ab_variants.py base= tensor=B200REG_ZM_TMA=2,lib=libb200reg_tma.so tensor_tx64=B200REG_ZM_TMA=2,B200REG_ZM_TX32=0,lib=libb200reg_tma.so \
    cp_async=B200REG_ZM_TMA=0,lib=libb200reg_tma.so cp_async_tx64=B200REG_ZM_TMA=0,B200REG_ZM_TX32=0,lib=libb200reg_tma.so \
    tensor_l2=B200REG_ZM_TMA=2,B200REG_ZM_TMA_L2=2,lib=libb200reg_tma.so rows=B200REG_ZM_TMA=1,lib=libb200reg_tma.so > gpurun_out/r02_ab_tma.log 2>&1
tail -6 gpurun_out/r02_ab_tma.log
python profiles/bench_extras.py > gpurun_out/r02_bench_extras.json 2> gpurun_out/r02_bench_extras.err
tail -3 gpurun_out/r02_bench_extras.json
