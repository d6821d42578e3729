#!/bin/bash
# GPU call 19 of round 2 (8 GPUs): host <-> device copy bandwidth per rank with 1 and with 8 ranks copying at once (the ceiling of e2e scaling)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 120 python profiles/exp_pcie_contention.py 2>&1 | grep PCIE | tee gpurun_out/r02s_pcie_contention.log
for N in 2 4 8; do
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N profiles/exp_pcie_contention.py 2>&1 | grep PCIE | tee -a gpurun_out/r02s_pcie_contention.log
done
nproc >> gpurun_out/r02s_pcie_contention.log; free -g | head -2 >> gpurun_out/r02s_pcie_contention.log
