"""Fast mode (float32 fields + float32 FMA smoothing inside the Demons loop) against parity mode on BASELINE configs[1]: device time
of the three level loops and of the whole registration, elapsed iterations, and the error of the fast displacement field against
the parity field (max, percentiles), in mm.  Prints one line: FAST {json}."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from platipy_b200 import registration as reg
from platipy_b200.engine import Engine
from platipy_b200.synth import synth_pair

size = tuple(int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (512, 512, 256)
eng = Engine.get(0)
fixed, moving = synth_pair(size, seed=0, moving_seed=100)
dF, dM = eng.to_device(fixed), eng.to_device(moving)
kw = dict(resolution_staging=[4, 2, 1], iteration_staging=[100, 50, 25])
out = {}
fields = {}
for mode in ("parity", "fast"):
    best = None
    for _ in range(3):
        eng.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(eng.stream)
        img, tfm, dvf = reg.fast_symmetric_forces_demons_registration(dF, dM, precision=mode, **kw)
        e1.record(eng.stream)
        eng.synchronize()
        st = reg.LAST_LEVEL_STATS
        total = e0.elapsed_time(e1)
        if best is None or total < best["registration_ms"]:
            best = {"registration_ms": total, "levels_ms": [s["gpu_ms"] for s in st], "elapsed": [s["elapsed_iterations"] for s in st],
                    "full_res_ms_per_iteration": st[-1]["gpu_ms"] / max(1, st[-1]["elapsed_iterations"]), "metric": [s["metric"] for s in st]}
    out[mode] = best
    fields[mode] = dvf.tensor.clone()
err = (fields["fast"] - fields["parity"]).abs().reshape(-1)
sub = err[:: max(1, err.numel() // 20_000_000)]
q = torch.quantile(sub.to(torch.float64)[:16_000_000], torch.tensor([0.5, 0.99, 0.999], dtype=torch.float64, device=sub.device))
out["error_mm"] = {"max": float(err.max()), "median": float(q[0]), "p99": float(q[1]), "p99.9": float(q[2]),
                   "field_abs_max_mm": float(fields["parity"].abs().max())}
out["speedup_full_res_iteration"] = out["parity"]["full_res_ms_per_iteration"] / out["fast"]["full_res_ms_per_iteration"]
out["speedup_registration"] = out["parity"]["registration_ms"] / out["fast"]["registration_ms"]
print("FAST " + json.dumps(out))
