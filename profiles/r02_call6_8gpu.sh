#!/bin/bash
# GPU call 6 of round 2 (8 GPUs): dist_check on 8 ranks (more ranks than atlases / structures), the bench line at N = 4 and N = 8.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tests/dist_check.py > gpurun_out/r02f_dist_check_8gpu.log 2>&1
grep -c "equal True, probabilities equal True" gpurun_out/r02f_dist_check_8gpu.log; grep -c "False" gpurun_out/r02f_dist_check_8gpu.log; grep -c ": OK" gpurun_out/r02f_dist_check_8gpu.log
for N in 4 8; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r02f_bench_${N}gpu.json 2> gpurun_out/r02f_bench_${N}gpu.err
python - $N <<'PY'
import json, sys
n = sys.argv[1]
d = json.loads([l for l in open(f"gpurun_out/r02f_bench_{n}gpu.json").read().strip().splitlines() if l.startswith("{")][-1])
print({k: d[k] for k in ("value", "ms_per_step", "n_gpus")}, d["e2e"]["ms_per_step"], d["e2e_single_call"]["ms_per_step"], d.get("numa_binding_rank0"))
for name, f in d["fusion"].items():
    print(name, f.get("wall_ms"), f.get("stages_ms"), f.get("dice_vs_truth_min"))
    print(name, "checksums", list(f.get("mask_checksums", {}).items())[:3])
PY
tail -2 gpurun_out/r02f_bench_${N}gpu.err
done
nvidia-smi topo -m > gpurun_out/r02f_topo.txt 2>&1; (cat /sys/bus/pci/devices/*/numa_node | sort | uniq -c) >> gpurun_out/r02f_topo.txt 2>&1; nproc >> gpurun_out/r02f_topo.txt
