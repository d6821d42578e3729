"""Informational: device time at 512x512x256 of the entry points added in round 1's third session (written without GPU time).
Every row is measured on its own and a failure is recorded as text for that row.  Prints one "EXP {json}" line."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from platipy_b200 import comparison, fusion, generation, label_utils as lu, linear
from platipy_b200.engine import Engine
from platipy_b200.sitk_compat import Image
from platipy_b200.synth import synth_labels, synth_pair

SIZE = (512, 512, 256)
eng = Engine.get(0)
fixed, moving = synth_pair(SIZE, seed=0, moving_seed=100)
dF, dM = eng.to_device(fixed), eng.to_device(moving)
labels = [eng.to_device(Image(l)) for l in synth_labels(SIZE, 2, seed=200)]
I3, Z3 = np.eye(3), np.zeros(3)


def timed(fn, reps=3):
    fn()
    eng.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(eng.stream)
    for _ in range(reps):
        fn()
    e1.record(eng.stream)
    eng.synchronize()
    return e0.elapsed_time(e1) / reps


def wall(fn):
    fn()
    eng.synchronize()
    t0 = time.perf_counter()
    fn()
    eng.synchronize()
    return 1e3 * (time.perf_counter() - t0)


rows = {}


def row(name, fn, how=timed):
    try:
        rows[name] = how(fn)
    except Exception as e:  # noqa: BLE001
        rows[name] = "error: " + repr(e)[:160]


row("signed_maurer_distance_map_ms", lambda: eng.signed_maurer_distance_map(labels[0]))
row("label_contour_ms", lambda: eng.label_contour(labels[0]))
row("binary_dilate_r3_ms", lambda: eng.binary_dilate(labels[0], lu.ball_offsets((3, 3, 3))))
row("image_moments_ms", lambda: eng.image_moments(dF))
row("linreg_correlation_fullres_stride4_ms", lambda: eng.linreg_correlation(dF, dM, I3, Z3, I3, Z3, None, None, 4))
try:
    f_bins, m_bins = linear.mattes_bins(*eng.minmax(dF)), linear.mattes_bins(*eng.minmax(dM))
    _, table, _ = linear.mattes_value_and_table(eng.linreg_mattes_histogram(dF, dM, I3, Z3, f_bins, m_bins, 50, None, None, 4)[0])
    row("linreg_mattes_histogram_fullres_stride4_ms", lambda: eng.linreg_mattes_histogram(dF, dM, I3, Z3, f_bins, m_bins, 50, None, None, 4))
    row("linreg_mattes_derivative_fullres_stride4_ms", lambda: eng.linreg_mattes_derivative(dF, dM, I3, Z3, I3, Z3, f_bins, m_bins, table, None, None, 4))
except Exception as e:  # noqa: BLE001
    rows["linreg_mattes"] = "error: " + repr(e)[:160]
row("patch_correlation_weight_map_3mm_window8_ms", lambda: fusion.compute_weight_map(dF, dM, "patch_correlation", fusion.DEFAULT_VOTE_PARAMS), wall)
row("compute_surface_metrics_wall_ms", lambda: comparison.compute_surface_metrics(labels[0], labels[1]), wall)
row("compute_metric_dsc_wall_ms", lambda: comparison.compute_metric_dsc(labels[0], labels[1]), wall)
row("generate_field_shift_wall_ms", lambda: generation.generate_field_shift(labels[0], (5, 5, 5), 3), wall)
rows["voxels"] = SIZE[0] * SIZE[1] * SIZE[2]
print("EXP " + json.dumps(rows))
