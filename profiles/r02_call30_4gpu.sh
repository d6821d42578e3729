#!/bin/bash
# GPU call 30 of round 2 (4 GPUs): the bench line of the final build at N = 2 and N = 4 (the driver's scaling run does N = 1, 2, 4, 8)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for N in 2 4; do
SECONDS=0
CUDA_VISIBLE_DEVICES=$(seq -s, 0 $((N-1))) timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$N bench.py --gpus $N --steps 6 --warmup 3 > gpurun_out/r02_final_bench_${N}gpu.json 2> gpurun_out/r02_final_bench_${N}gpu.err
echo "N=$N bench wall ${SECONDS}s"
python - $N <<'PY'
import json, sys
n = sys.argv[1]
d = json.loads([l for l in open(f"gpurun_out/r02_final_bench_{n}gpu.json").read().strip().splitlines() if l.startswith("{")][-1])
print({k: d[k] for k in ("value", "ms_per_step", "n_gpus")}, "e2e", d["e2e"]["ms_per_step"], "single", d["e2e_single_call"]["ms_per_step"])
for name, f in d["fusion"].items():
    print(name, f.get("wall_ms"), f.get("dice_vs_truth_min"))
PY
done
