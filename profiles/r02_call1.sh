#!/bin/bash
# GPU call 1 of round 2: smoke, the whole GPU suite, A/B of the tensor-map staging (now the default) against cp.async, the bench line
# with the fusion workloads.   gpurun --timeout 1500 -- 'bash profiles/r02_call1.sh'
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
(nvidia-smi -L; nproc; free -g | head -2; numactl -H 2>/dev/null | head -5) > gpurun_out/r02a_box.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02a_smoke.log 2>&1
tail -2 gpurun_out/r02a_smoke.log
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 300 > gpurun_out/r02a_pytest_gpu.log 2>&1
tail -5 gpurun_out/r02a_pytest_gpu.log
timeout 300 python profiles/ab_variants.py tma_l2_128= cp_async=B200REG_ZM_TMA=0 tma_l2_none=B200REG_ZM_TMA_L2=0 tma_l2_256=B200REG_ZM_TMA_L2=3 > gpurun_out/r02a_ab_tma.log 2>&1
tail -5 gpurun_out/r02a_ab_tma.log
timeout 700 python bench.py --steps 5 --warmup 3 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
tail -c 3000 gpurun_out/r02a_bench.json; tail -5 gpurun_out/r02a_bench.err
