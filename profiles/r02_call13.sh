#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 900 -x > gpurun_out/r02m_pytest_gpu.log 2>&1
tail -3 gpurun_out/r02m_pytest_gpu.log
timeout 200 python profiles/exp_registration_total.py 2>&1 | grep TOTAL | tee gpurun_out/r02m_registration_total.log
