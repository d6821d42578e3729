"""A/B of kernel variants on one box: every configuration (environment switches and/or an alternative build of the
library) runs BASELINE configs[1] in its own process; reports the device time of the full-resolution level per
elapsed iteration, the whole registration, and whether the DVF (sub-sampled) and the per-level metrics are
bit-identical to the first configuration.

    python profiles/ab_variants.py [name=ENV1=v,ENV2=v[,lib=libb200reg_x.so]] ...

Without arguments: the built-in list below.  The environment switches that select rejected kernel variants only act on a
library built with them: `make -C platipy_b200/csrc OUT=../libb200reg_ab.so EXTRA=-DB200REG_AB_VARIANTS` and `lib=libb200reg_ab.so`."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, os, json, numpy as np
sys.path.insert(0, %r)
from platipy_b200 import registration as reg
from platipy_b200.engine import Engine
from platipy_b200.synth import synth_pair
eng = Engine.get(0)
from platipy_b200 import sitk_compat as sk
if os.path.exists("/tmp/ab_fixed.npy"):
    f, m = sk.Image(np.load("/tmp/ab_fixed.npy")), sk.Image(np.load("/tmp/ab_moving.npy"))
else:
    f, m = synth_pair((512, 512, 256), seed=0, moving_seed=100)
    np.save("/tmp/ab_fixed.npy", f.array); np.save("/tmp/ab_moving.npy", m.array)
dF, dM = eng.to_device(f), eng.to_device(m)
best = None
for _ in range(3):
    img, tfm, dvf = reg.fast_symmetric_forces_demons_registration(dF, dM, resolution_staging=[4, 2, 1], iteration_staging=[100, 50, 25])
    st = reg.LAST_LEVEL_STATS
    tot = sum(s["gpu_ms"] for s in st)
    if best is None or tot < best[0]:
        best = (tot, [(s["elapsed_iterations"], s["gpu_ms"], s["metric"], s["rms_change"]) for s in st])
np.save(sys.argv[1], eng.to_host(dvf, pinned=False).array[::4, ::4, ::4].copy())
print("STATS " + json.dumps({"levels_ms": best[0], "levels": best[1]}))
''' % ROOT

DEFAULT = [
    "baseline=B200REG_UPDATE_BRANCHY=1,B200REG_ZM_REGADD=0",
    "update_straight=B200REG_ZM_REGADD=0",
    "zm_regadd=B200REG_UPDATE_BRANCHY=1",
    "both=",
]


def main():
    specs = sys.argv[1:] or DEFAULT
    out, ref = {}, None
    for spec in specs:
        name, _, rest = spec.partition("=")
        env = dict(os.environ)
        for kv in [x for x in rest.split(",") if x]:
            k, _, v = kv.partition("=")
            if k == "lib":
                env["B200REG_LIB"] = os.path.join(ROOT, "platipy_b200", v)
            else:
                env[k] = v
        path = f"/tmp/ab_{name}.npy"
        r = subprocess.run([sys.executable, "-c", CHILD, path], env=env, capture_output=True, text=True)
        line = [l for l in r.stdout.splitlines() if l.startswith("STATS")]
        if not line:
            out[name] = {"error": r.stderr[-800:]}
            continue
        st = json.loads(line[0][6:])
        full = st["levels"][-1]
        res = {"full_res_ms_per_iteration": full[1] / max(full[0], 1), "levels_ms": st["levels_ms"], "elapsed": [l[0] for l in st["levels"]]}
        dvf = np.load(path)
        if ref is None:
            ref = (dvf, st["levels"])
        res["dvf_identical_to_first"] = bool(np.array_equal(dvf, ref[0]))
        res["max_abs_dvf_diff_mm"] = float(np.abs(dvf - ref[0]).max())
        res["metrics_identical_to_first"] = [l[2:] for l in st["levels"]] == [l[2:] for l in ref[1]]
        out[name] = res
        print(name, json.dumps(res), flush=True)
    print("AB " + json.dumps(out))


if __name__ == "__main__":
    main()
