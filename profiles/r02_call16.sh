#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 900 -x > gpurun_out/r02p_pytest_gpu.log 2>&1
tail -3 gpurun_out/r02p_pytest_gpu.log
for v in on_grid_off on_grid_on rs3 on_grid_off on_grid_on rs3; do
  unset PLATIPY_B200_LIB
  if [ $v = on_grid_off ]; then export PLATIPY_B200_IDENTITY_COPY=0; else unset PLATIPY_B200_IDENTITY_COPY; fi
  if [ $v = rs3 ]; then export PLATIPY_B200_LIB=$PWD/platipy_b200/libb200reg_rs3.so; fi
  echo "$v $(timeout 200 python profiles/exp_registration_total.py 2>&1 | grep TOTAL)" | tee -a gpurun_out/r02p_ab_on_grid.log
done
unset PLATIPY_B200_LIB PLATIPY_B200_IDENTITY_COPY
timeout 300 python bench.py --steps 3 --warmup 2 --no-fusion --no-fast-mode --no-cpu-baseline > gpurun_out/r02p_bench.json 2>/dev/null
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02p_bench.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["e2e"]["ms_per_step"], d["resample_cfg3"]["batched"]["ms"], d["resample_cfg3"]["per_call"]["ms"])
PY
