#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 900 -x > gpurun_out/r02q_pytest_gpu.log 2>&1
tail -3 gpurun_out/r02q_pytest_gpu.log
for v in ckpt_off ckpt_on ckpt_off ckpt_on; do
  if [ $v = ckpt_off ]; then export PLATIPY_B200_DERICHE_CKPT=0; else unset PLATIPY_B200_DERICHE_CKPT; fi
  echo "$v $(timeout 200 python profiles/exp_registration_total.py 2>&1 | grep TOTAL)" | tee -a gpurun_out/r02q_ab_deriche_ckpt.log
done
unset PLATIPY_B200_DERICHE_CKPT
timeout 300 python bench.py --steps 3 --warmup 2 --no-fusion --no-fast-mode --no-cpu-baseline > gpurun_out/r02q_bench.json 2>/dev/null
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02q_bench.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["e2e"]["ms_per_step"], d["resample_cfg3"]["batched"]["ms"], d["resample_cfg3"]["per_call"]["ms"])
PY
