#!/bin/bash
# GPU call 2 of round 2 (1 GPU): the whole GPU suite (new: whole cfg2 parity, semantic switches, iteration events, pipelined host API,
# bit-packed label propagation), the bench line, the launch list of one registration and full ncu captures of the iteration kernels.
#   gpurun --timeout 1800 -- 'bash profiles/r02_call2.sh'
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 -s > gpurun_out/r02b_pytest_gpu.log 2>&1
tail -5 gpurun_out/r02b_pytest_gpu.log; grep "cfg2 whole" gpurun_out/r02b_pytest_gpu.log
timeout 700 python bench.py --steps 5 --warmup 3 > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err
tail -c 1500 gpurun_out/r02b_bench.json; tail -3 gpurun_out/r02b_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02b_launches_registration.csv python profiles/prof_registration.py > gpurun_out/r02b_prof.log 2>&1
tail -2 gpurun_out/r02b_prof.log
# full captures: the four kernels of a full-resolution iteration (launches well inside level 2) -- 3 launches each
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'conv3d_zm2_kernel|demons_warp2_kernel|demons_force2_kernel' -s 620 -c 12 -o gpurun_out/r02b_iteration_kernels python profiles/prof_registration.py > gpurun_out/r02b_ncu.log 2>&1
tail -3 gpurun_out/r02b_ncu.log
ls -la gpurun_out/*.ncu-rep
