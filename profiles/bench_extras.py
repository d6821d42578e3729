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
print(json.dumps({"apply_transform_21_calls_ms": ms_calls, "apply_transform_21_calls_GBs": bytes_calls / ms_calls / 1e6,
                  "apply_transform_batched_ms": ms_batch, "apply_transform_batched_GBs": bytes_batch / ms_batch / 1e6,
                  "voxels_per_s_batched": 21 * N / (ms_batch * 1e-3),
                  "combine_labels_8_atlases_1_structure_ms": ms_vote, "staple_8_raters_ms": ms_staple, "staple_iterations": info["elapsed_iterations"]}))
