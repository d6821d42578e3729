#!/bin/bash
# GPU call 7 of round 2 (1 GPU): compute-sanitizer over the hot path, the GPU suite on the current build, a short bench.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
bash profiles/r02_sanitize.sh
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 > gpurun_out/r02g_pytest_gpu.log 2>&1
tail -4 gpurun_out/r02g_pytest_gpu.log
timeout 600 python bench.py --steps 6 --warmup 3 --no-fusion --no-cfg3 --no-cpu-baseline > gpurun_out/r02g_bench.json 2> gpurun_out/r02g_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02g_bench.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step")}, d["e2e"]["ms_per_step"], d["e2e_single_call"]["ms_per_step"], d["roofline"]["ms_per_launch"], d["clocks"])
PY
tail -3 gpurun_out/r02g_bench.err
