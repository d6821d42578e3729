"""Small driver for ncu: a few full-resolution Demons iterations (512x512x256) and one batched label warp.
Usage (under gpurun): ncu ... python profiles/prof_demons.py [iters]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from platipy_b200 import registration as reg
from platipy_b200.engine import Engine
from platipy_b200.synth import synth_pair

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 3
size = tuple(int(v) for v in sys.argv[2:5]) if len(sys.argv) >= 5 else (512, 512, 256)
eng = Engine.get(0)
fixed, moving = synth_pair(size, seed=0, moving_seed=100)
dF, dM = eng.to_device(fixed), eng.to_device(moving)
f = reg.FastSymmetricForcesDemonsRegistrationFilter()
f.SetStandardDeviations((1.5, 1.5, 1.5))
f.SetSmoothUpdateField(True)
dvf, st = eng.demons_execute(dF, dM, f.params(iters))
eng.synchronize()
print(st)
