#!/bin/bash
# GPU call 18 of round 2 (8 GPUs): the whole bench line at N = 8 with the pinned-buffer pool (e2e) and the final kernels; what the driver's scaling run executes.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=8
SECONDS=0
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r02r_bench_${N}gpu.json 2> gpurun_out/r02r_bench_${N}gpu.err
echo "bench wall ${SECONDS}s"
python - $N <<'PY'
import json, sys
n = sys.argv[1]
d = json.loads([l for l in open(f"gpurun_out/r02r_bench_{n}gpu.json").read().strip().splitlines() if l.startswith("{")][-1])
print({k: d[k] for k in ("value", "ms_per_step", "n_gpus")}, d["e2e"]["ms_per_step"], d["e2e_single_call"]["ms_per_step"], d.get("numa_binding_rank0"))
for name, f in d["fusion"].items():
    print(name, f.get("wall_ms"), f.get("stages_ms"), f.get("dice_vs_truth_min"))
PY
tail -2 gpurun_out/r02r_bench_${N}gpu.err
