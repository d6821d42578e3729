"""Device time of the WHOLE registration (pyramid + level loops + glue + final warp), best of 5, and of the level loops alone, for the library
named by PLATIPY_B200_LIB / B200REG_LIB (default: the in-tree build).  Prints: TOTAL {json}."""
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
out = {}
for name, kw in (("headline [4,2,1]x[100,50,25]", dict(resolution_staging=[4, 2, 1], iteration_staging=[100, 50, 25])),
                 ("platipy default [8,4,1]x[10,10,10]", dict(resolution_staging=[8, 4, 1], iteration_staging=[10, 10, 10]))):
    best = None
    for _ in range(6):
        eng.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(eng.stream)
        reg.fast_symmetric_forces_demons_registration(dF, dM, **kw)
        e1.record(eng.stream)
        eng.synchronize()
        t = e0.elapsed_time(e1)
        loops = sum(s["gpu_ms"] for s in reg.LAST_LEVEL_STATS)
        if best is None or t < best[0]:
            best = (t, loops)
    out[name] = {"registration_ms": best[0], "level_loops_ms": best[1], "glue_ms": best[0] - best[1]}
print("TOTAL " + json.dumps(out))
