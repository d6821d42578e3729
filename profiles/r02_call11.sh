#!/bin/bash
# GPU call 11 of round 2 (1 GPU): the whole GPU suite incl. the cfg4-size pipeline parity test, on the current build.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 900 --durations=8 > gpurun_out/r02k_pytest_gpu.log 2>&1
tail -16 gpurun_out/r02k_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
