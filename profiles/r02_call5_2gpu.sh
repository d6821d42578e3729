#!/bin/bash
# GPU call 5 of round 2 (2 GPUs): the sharded pipeline against the single-process run (tests/dist_check.py), then the bench line at N = 2.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_check.py > gpurun_out/r02e_dist_check_2gpu.log 2>&1
grep -c "equal True, probabilities equal True" gpurun_out/r02e_dist_check_2gpu.log; grep -c "False" gpurun_out/r02e_dist_check_2gpu.log; tail -3 gpurun_out/r02e_dist_check_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02e_bench_2gpu.json 2> gpurun_out/r02e_bench_2gpu.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r02e_bench_2gpu.json").read().strip().splitlines() if l.startswith("{")][-1])
print({k: d[k] for k in ("value", "ms_per_step", "n_gpus", "e2e", "e2e_single_call")})
print(d.get("numa_binding_rank0"))
for n, f in d["fusion"].items():
    print(n, f.get("wall_ms"), f.get("stages_ms"), f.get("dice_vs_truth_min"), f.get("exchange_payload_bytes_per_rank"))
    print(n, "checksums", list(f.get("mask_checksums", {}).items())[:3])
PY
tail -5 gpurun_out/r02e_bench_2gpu.err
