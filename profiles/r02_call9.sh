#!/bin/bash
# GPU call 9 of round 2 (1 GPU): the bench exactly as the driver runs it (default flags, then 20 / 5), the reference arm, the suite.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo skip-suite

time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02i_bench.json 2> gpurun_out/r02i_bench.err
tail -4 gpurun_out/r02i_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02i_bench.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step")}, d["e2e"]["ms_per_step"], d["e2e_single_call"]["ms_per_step"], d["roofline"]["ms_per_launch"], d["roofline"]["frac"], d["clocks"])
print(d["cpu_baseline"])
print({k: v for k, v in d["fast_mode"].items() if k != "error_vs_parity_mm"}, d["fast_mode"].get("error_vs_parity_mm"))
for n, f in d["fusion"].items():
    print(n, f.get("wall_ms"), f.get("error"))
print(d["resample_cfg3"]["batched"]["ms"], d["resample_cfg3"]["per_call"]["ms"])
PY
time timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02i_bench_reference.json 2> gpurun_out/r02i_bench_reference.err
tail -4 gpurun_out/r02i_bench_reference.err; cut -c1-900 gpurun_out/r02i_bench_reference.json
