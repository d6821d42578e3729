"""Secondary measurements (not the headline): BASELINE.json configs[2] (apply_transform of one CT + 20 masks through a
dense DVF, per-call and batched) and the fusion kernels at 512x512x256.  Prints one JSON line."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from platipy_b200 import registration as reg
from platipy_b200 import sitk_compat as sk
from platipy_b200.engine import Engine
from platipy_b200.sitk_compat import Image
from platipy_b200.synth import smooth_random_dvf, synth_labels, synth_pair

SIZE = (512, 512, 256)
N = SIZE[0] * SIZE[1] * SIZE[2]
eng = Engine.get(0)
fixed, moving = synth_pair(SIZE, seed=0, moving_seed=100)
labels = [eng.to_device(Image(l)) for l in synth_labels(SIZE, 20, seed=200)]
dM, dF = eng.to_device(moving), eng.to_device(fixed)
tfm = sk.DisplacementFieldTransform(Image(smooth_random_dvf(SIZE, seed=9, peak_mm=6.0), is_vector=True))
imgs = [dM] + labels
dvs, ips = [-1000] + [0] * 20, [sk.sitkLinear] + [sk.sitkNearestNeighbor] * 20


def timed(fn, reps=5):
    fn()
    eng.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(eng.stream)
    for _ in range(reps):
        fn()
    e1.record(eng.stream)
    eng.synchronize()
    return e0.elapsed_time(e1) / reps


def per_call():
    for im, dv, ip in zip(imgs, dvs, ips):
        reg.apply_transform(im, dF, tfm, dv, ip)


ms_calls = timed(per_call)
ms_batch = timed(lambda: reg.apply_transform_batch(imgs, dF, tfm, dvs, ips))
bytes_calls = N * (32 + 20 * 26)   # f32 linear 32 B/voxel, u8 NN 26 B/voxel (f64 DVF read per call)
bytes_batch = N * (24 + 8 + 40)    # DVF once + f32 in/out + 20 x (1 + 1)
# fusion
w = eng.weight_map(dF, dM, 0)
num, den = eng.empty(labels[0].tensor.shape, np.float32), eng.empty(labels[0].tensor.shape, np.float32)


def vote8():
    for a in range(8):
        eng.vote_accumulate(labels[a], w, num, den, a == 0)
    eng.vote_finalize(num, den, dF, 1.0, 1e-4)


ms_vote = timed(vote8, 3)
rng_labels = labels[:8]
import time

t0 = time.perf_counter()
_, info = eng.staple(rng_labels, threshold=1e-4, rescale=True)
eng.synchronize()
ms_staple = 1e3 * (time.perf_counter() - t0)
# ---- rows added in the second session (512 x 512 x 256 each) --------------------------------------------------------
from platipy_b200 import comparison, fusion, generation, label_utils as lu, linear

prob = eng.vote_finalize(num, den, dF, 1.0, 1e-4)
ms_ppi = timed(lambda: eng.process_probability(prob, 0.5), 3)
ms_fill = timed(lambda: eng.binary_fillhole(labels[0]), 3)
ms_cc = timed(lambda: eng.largest_component(labels[0]), 3)
ms_block = timed(lambda: eng.weight_map_block(dF, dM, (5, 5, 5), 1e12, 6), 3)
ms_bspline = timed(lambda: reg.apply_transform(dM, dF, tfm, -1000, sk.sitkBSpline), 3)
ms_closing = timed(lambda: eng.binary_closing(labels[0], (3, 3, 3), lu.ball_offsets((3, 3, 3))), 2)
ms_overlap = timed(lambda: eng.resolve_overlap(labels[:5]), 3)
t0 = time.perf_counter()
_, ltfm = linear.linear_registration(dF, dM, reg_method="similarity", shrink_factors=[8, 2, 1], smooth_sigmas=[4, 2, 0], default_value=-1000)
eng.synchronize()
ms_linear = 1e3 * (time.perf_counter() - t0)
n_eval = sum(len(h) for h in linear.LAST_HISTORY)
ms_metric = timed(lambda: eng.linreg_meansq(dF, dM, np.eye(3), np.zeros(3), np.eye(3), np.zeros(3), None, None, 4), 10)
# rows added in the third session of round 1 (not yet measured on a GPU: the round's budget was spent)
f_bins, m_bins = linear.mattes_bins(*eng.minmax(dF)), linear.mattes_bins(*eng.minmax(dM))
_, table, _ = linear.mattes_value_and_table(eng.linreg_mattes_histogram(dF, dM, np.eye(3), np.zeros(3), f_bins, m_bins, 50, None, None, 4)[0])
session3 = {
    "signed_maurer_distance_map_ms": timed(lambda: eng.signed_maurer_distance_map(labels[0]), 3),
    "label_contour_ms": timed(lambda: eng.label_contour(labels[0]), 3),
    "binary_dilate_r3_ms": timed(lambda: eng.binary_dilate(labels[0], lu.ball_offsets((3, 3, 3))), 2),
    "patch_correlation_3mm_w8_ms": timed(lambda: fusion.compute_weight_map(dF, dM, "patch_correlation", fusion.DEFAULT_VOTE_PARAMS), 2),
    "linreg_correlation_fullres_stride4_ms": timed(lambda: eng.linreg_correlation(dF, dM, np.eye(3), np.zeros(3), np.eye(3), np.zeros(3), None, None, 4), 10),
    "linreg_mattes_histogram_fullres_stride4_ms": timed(lambda: eng.linreg_mattes_histogram(dF, dM, np.eye(3), np.zeros(3), f_bins, m_bins, 50, None, None, 4), 10),
    "linreg_mattes_derivative_fullres_stride4_ms": timed(
        lambda: eng.linreg_mattes_derivative(dF, dM, np.eye(3), np.zeros(3), np.eye(3), np.zeros(3), f_bins, m_bins, table, None, None, 4), 10),
    "image_moments_ms": timed(lambda: eng.image_moments(dF), 5),
}
t0 = time.perf_counter()
comparison.compute_surface_metrics(labels[0], labels[1])
session3["compute_surface_metrics_wall_ms"] = 1e3 * (time.perf_counter() - t0)
t0 = time.perf_counter()
generation.generate_field_shift(labels[0], (5, 5, 5), 3)
eng.synchronize()
session3["generate_field_shift_wall_ms"] = 1e3 * (time.perf_counter() - t0)
print(json.dumps(session3))
extra = {"process_probability_image_ms": ms_ppi, "binary_fillhole_ms": ms_fill, "largest_component_ms": ms_cc,
         "weight_map_block_r5_ms": ms_block, "apply_transform_bspline_f32_ms": ms_bspline, "binary_closing_r3_ms": ms_closing,
         "correct_volume_overlap_5_ms": ms_overlap, "linear_registration_similarity_8_2_1_ms": ms_linear,
         "linear_registration_iterations": n_eval, "linreg_meansq_fullres_stride4_ms": ms_metric,
         "linreg_meansq_fullres_Gsamples_per_s": N / 4 / (ms_metric * 1e-3) / 1e9}
print(json.dumps(extra))
print(json.dumps({"apply_transform_21_calls_ms": ms_calls, "apply_transform_21_calls_GBs": bytes_calls / ms_calls / 1e6,
                  "apply_transform_batched_ms": ms_batch, "apply_transform_batched_GBs": bytes_batch / ms_batch / 1e6,
                  "voxels_per_s_batched": 21 * N / (ms_batch * 1e-3),
                  "combine_labels_8_atlases_1_structure_ms": ms_vote, "staple_8_raters_ms": ms_staple, "staple_iterations": info["elapsed_iterations"]}))
