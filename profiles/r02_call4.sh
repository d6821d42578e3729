#!/bin/bash
# GPU call 4 of round 2 (1 GPU): programmatic dependent launch on / off, the pipelined host API after the read-back fix, bench, suite.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python profiles/ab_variants.py pdl_on= pdl_off=B200REG_PDL=0 pdl_on_again= > gpurun_out/r02d_ab_pdl.log 2>&1
grep -v "^AB" gpurun_out/r02d_ab_pdl.log | cut -c1-330
timeout 200 python profiles/exp_pipeline_timeline.py > gpurun_out/r02d_timeline.log 2>&1
tail -c 3500 gpurun_out/r02d_timeline.log
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 > gpurun_out/r02d_pytest_gpu.log 2>&1
tail -4 gpurun_out/r02d_pytest_gpu.log
timeout 700 python bench.py --steps 6 --warmup 3 > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02d_bench.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "e2e", "e2e_single_call")})
print(d["roofline"]["ms_per_launch"], d["roofline"]["frac"], d["clocks"], [round(l["gpu_ms"], 2) for l in d["levels"]])
for n, f in d["fusion"].items():
    print(n, f.get("wall_ms"), f.get("stages_ms"))
print(d["resample_cfg3"]["batched"])
PY
tail -3 gpurun_out/r02d_bench.err
