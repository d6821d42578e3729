#!/bin/bash
# GPU call 24 of round 2: ncu --set full of the glue kernels of a full-resolution registration (field re-gridding / composition, recursive Gaussian,
# min / max); the report is summarised on the box (profiles/ncu_summary.py) because it exceeds what gpurun copies back
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
timeout 900 ncu --set full --clock-control none -k regex:'resample_vec3_kernel|deriche_|minmax_partial' \
  -o /tmp/r02x_glue_kernels -f python profiles/prof_registration.py > gpurun_out/r02x_ncu.log 2>&1
tail -2 gpurun_out/r02x_ncu.log
python profiles/ncu_summary.py /tmp/r02x_glue_kernels.ncu-rep > gpurun_out/r02x_ncu_glue_kernels_summary.txt 2>&1
wc -l gpurun_out/r02x_ncu_glue_kernels_summary.txt
