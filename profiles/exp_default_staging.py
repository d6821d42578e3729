"""Informational: the headline volume (512x512x256) registered with platipy's OWN default arguments -- shrink factors [8, 4, 1],
10 iterations per level, smoothing sigmas [8, 4, 1] mm (deformable.py:190-204) -- device resident, CUDA-event time.  SURVEY 8d asks for
this number next to BASELINE.json configs[1].  Prints one "EXP {json}" line."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from platipy_b200 import registration as reg
from platipy_b200.engine import Engine
from platipy_b200.synth import synth_pair

eng = Engine.get(0)
fixed, moving = synth_pair((512, 512, 256), seed=0, moving_seed=100)
dF, dM = eng.to_device(fixed), eng.to_device(moving)
reg.fast_symmetric_forces_demons_registration(dF, dM)
eng.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
steps = 3
e0.record(eng.stream)
for _ in range(steps):
    reg.fast_symmetric_forces_demons_registration(dF, dM)
e1.record(eng.stream)
eng.synchronize()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
st = reg.LAST_LEVEL_STATS
vox_it = float(sum(s["voxels"] * s["elapsed_iterations"] for s in st))
print("EXP " + json.dumps({"resolution_staging": [8, 4, 1], "iteration_staging": [10, 10, 10], "ms_per_registration": ms,
                           "value": vox_it / (ms * 1e-3) / 1e6, "unit": "Mvoxel*it/s", "elapsed_iterations": [s["elapsed_iterations"] for s in st],
                           "levels_gpu_ms": [s["gpu_ms"] for s in st]}))
