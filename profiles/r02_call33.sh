#!/bin/bash
# GPU call 33 of round 2: single nearest-neighbour labels through a field on the output grid take a four-rows-per-thread kernel -- tests + cfg3 per call A/B
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_multiatlas.py -m gpu -q -p no:cacheprovider --timeout 900 -x > gpurun_out/r02ad_pytest_gpu.log 2>&1
tail -3 gpurun_out/r02ad_pytest_gpu.log
for v in off on off on; do
  if [ $v = off ]; then export PLATIPY_B200_WARP_RESAMPLE=0; else unset PLATIPY_B200_WARP_RESAMPLE; fi
  timeout 300 python bench.py --steps 2 --warmup 1 --no-fusion --no-fast-mode --no-cpu-baseline > gpurun_out/r02ad_bench_$v.json 2>/dev/null
  python - $v <<'PY' | tee -a gpurun_out/r02ad_ab_nn_route.log
import json, sys
d = json.loads(open(f"gpurun_out/r02ad_bench_{sys.argv[1]}.json").read().strip().splitlines()[-1])
print("warp_resample", sys.argv[1], "cfg3 batched", d["resample_cfg3"]["batched"]["ms"], "per call", d["resample_cfg3"]["per_call"]["ms"], "step", d["ms_per_step"])
PY
done
