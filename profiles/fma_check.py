"""A/B of the strict build (-fmad=false, bit-exact vs the oracle) and an FMA-contracted build of the same kernels:
runs BASELINE configs[1] with each library in a subprocess, reports time per full-resolution iteration and the
maximum DVF difference between the two results (mm)."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, os, numpy as np
sys.path.insert(0, %r)
from platipy_b200 import registration as reg
from platipy_b200.engine import Engine
from platipy_b200.synth import synth_pair
eng = Engine.get(0)
f, m = synth_pair((512, 512, 256), seed=0, moving_seed=100)
dF, dM = eng.to_device(f), eng.to_device(m)
for _ in range(2):
    img, tfm, dvf = reg.fast_symmetric_forces_demons_registration(dF, dM, resolution_staging=[4, 2, 1], iteration_staging=[100, 50, 25])
st = reg.LAST_LEVEL_STATS
np.save(sys.argv[1], eng.to_host(dvf, pinned=False).array[::4, ::4, ::4].copy())
print("STATS", [(s["elapsed_iterations"], s["gpu_ms"]) for s in st])
''' % ROOT
out = {}
for name, lib in (("strict", "libb200reg.so"), ("fma", "libb200reg_fma.so")):
    env = dict(os.environ, B200REG_LIB=os.path.join(ROOT, "platipy_b200", lib))
    path = f"/tmp/dvf_{name}.npy"
    r = subprocess.run([sys.executable, "-c", CHILD, path], env=env, capture_output=True, text=True)
    line = [l for l in r.stdout.splitlines() if l.startswith("STATS")]
    out[name] = line[0] if line else r.stderr[-500:]
a, b = np.load("/tmp/dvf_strict.npy"), np.load("/tmp/dvf_fma.npy")
out["max_abs_dvf_difference_mm"] = float(np.abs(a - b).max())
print(json.dumps(out))
