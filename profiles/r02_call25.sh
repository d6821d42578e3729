#!/bin/bash
# GPU call 25 of round 2: block-shared scan-line row ends in the resampling kernels, four loads in flight in min / max -- GPU suite + A/B
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 900 -x > gpurun_out/r02y_pytest_gpu.log 2>&1
tail -4 gpurun_out/r02y_pytest_gpu.log
for v in per_voxel block per_voxel block; do
  if [ $v = per_voxel ]; then export PLATIPY_B200_LIB=$PWD/platipy_b200/libb200reg_noblockscan.so; else unset PLATIPY_B200_LIB; fi
  echo "scanline_$v $(timeout 200 python profiles/exp_registration_total.py 2>&1 | grep TOTAL)" | tee -a gpurun_out/r02y_ab_block_scanline.log
done
unset PLATIPY_B200_LIB
timeout 300 python bench.py --steps 3 --warmup 2 --no-fusion --no-fast-mode --no-cpu-baseline > gpurun_out/r02y_bench.json 2>/dev/null
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02y_bench.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["e2e"]["ms_per_step"], d["resample_cfg3"]["batched"]["ms"], d["resample_cfg3"]["per_call"]["ms"])
PY
