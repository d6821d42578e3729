#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cat > /tmp/dbg_fast.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
from platipy_b200 import registration as reg
from platipy_b200.engine import Engine
from platipy_b200.synth import synth_pair
eng = Engine.get(0)
fixed, moving = synth_pair((64, 48, 32), seed=0, moving_seed=100)
dF, dM = eng.to_device(fixed), eng.to_device(moving)
img, tfm, dvf = reg.fast_symmetric_forces_demons_registration(dF, dM, precision="fast", resolution_staging=[1], iteration_staging=[2])
eng.synchronize()
print("FAST OK", float(dvf.tensor.abs().max()))
PY
timeout 300 /usr/local/cuda/bin/compute-sanitizer --tool memcheck --print-limit 5 python /tmp/dbg_fast.py > gpurun_out/dbg_fast_memcheck.log 2>&1
grep -v "^=========     at\|^=========         Host\|^=========                in" gpurun_out/dbg_fast_memcheck.log | head -40
