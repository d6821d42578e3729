#!/bin/bash
# GPU call 3 of round 2 (1 GPU): depth of the TMA plane pipeline of the fused smoothing kernel (2 / 3 / 4 / 6 staged planes, cp.async),
# the timeline of the pipelined host API, the GPU suite on the current build.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 400 python profiles/ab_variants.py tma_4_stages= tma_2_stages=lib=libb200reg_st2.so tma_3_stages=lib=libb200reg_st3.so tma_6_stages=lib=libb200reg_st6.so cp_async=B200REG_ZM_TMA=0 > gpurun_out/r02c_ab_tma_stages.log 2>&1
grep -v "^AB" gpurun_out/r02c_ab_tma_stages.log | cut -c1-200
timeout 200 python profiles/exp_pipeline_timeline.py > gpurun_out/r02c_timeline.log 2>&1
tail -c 6000 gpurun_out/r02c_timeline.log
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 > gpurun_out/r02c_pytest_gpu.log 2>&1
tail -4 gpurun_out/r02c_pytest_gpu.log
