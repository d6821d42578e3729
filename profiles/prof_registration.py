"""ncu driver: one full fast_symmetric_forces_demons_registration (BASELINE configs[1]) on device-resident inputs."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from platipy_b200 import registration as reg
from platipy_b200.engine import Engine
from platipy_b200.synth import synth_pair

size = tuple(int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (512, 512, 256)
eng = Engine.get(0)
fixed, moving = synth_pair(size, seed=0, moving_seed=100)
dF, dM = eng.to_device(fixed), eng.to_device(moving)
img, tfm, dvf = reg.fast_symmetric_forces_demons_registration(dF, dM, resolution_staging=[4, 2, 1], iteration_staging=[100, 50, 25])
eng.synchronize()
print([{k: v for k, v in s.items() if k != "trace"} for s in reg.LAST_LEVEL_STATS])
