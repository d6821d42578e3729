#!/bin/bash
# GPU call 27 of round 2: ncu --set full of the pyramid Gaussian passes (tiled Float32 kernels) and of the restricted-level kernels
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none -k regex:'conv_x_f32_tiled|conv_yz_f32_tiled|conv_sel_f32|shrink_gather' -c 24 \
  -o /tmp/r02aa_pyramid_kernels -f python profiles/prof_registration.py 512 512 256 > gpurun_out/r02aa_ncu.log 2>&1
tail -2 gpurun_out/r02aa_ncu.log
python profiles/ncu_summary.py /tmp/r02aa_pyramid_kernels.ncu-rep > gpurun_out/r02aa_ncu_pyramid_kernels_summary.txt 2>&1
wc -l gpurun_out/r02aa_ncu_pyramid_kernels_summary.txt
