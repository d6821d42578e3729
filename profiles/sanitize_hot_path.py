"""Small end-to-end run of the hot path for compute-sanitizer (memcheck / racecheck / synccheck): a 3-level Demons registration at 64x56x40
(TMA-staged smoothing, PDL launches, device-resident halt), a restricted pyramid level, batched label propagation (bit-packed path), weighted vote, packed STAPLE,
process_probability_image (union-find CCL, fill-hole) and a short linear registration with the Mattes metric (histogram atomics)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from platipy_b200 import fusion, linear
from platipy_b200 import registration as reg
from platipy_b200 import sitk_compat as sk
from platipy_b200.engine import Engine
from platipy_b200.sitk_compat import Image
from platipy_b200.synth import synth_labels, synth_pair

eng = Engine.get(0)
size, sp = (64, 56, 40), (1.0, 1.0, 1.5)
fixed, moving = synth_pair(size, seed=1, spacing=sp, peak_mm=3.0)
img, tfm, dvf = reg.fast_symmetric_forces_demons_registration(fixed, moving, resolution_staging=[4, 2, 1], iteration_staging=[6, 4, 3])
# a level that shrinks by 8: blurred only where the level reads it (pyramid.cuh); the registered image and the level-start warp take the
# Demons warp kernel (resample_f32_on_grid_dvf)
reg.fast_symmetric_forces_demons_registration(fixed, moving, resolution_staging=[8, 2], iteration_staging=[3, 2])
reg.smooth_and_resample(fixed, shrink_factor=[8, 4, 3], smoothing_sigma=[4.0, 2.0, 1.5])
labels = [Image(l, sp) for l in synth_labels(size, 8, seed=300)]
outs = reg.apply_transform_batch([moving] + labels, fixed, tfm, [-1000] + [0] * 8, [sk.sitkLinear] + [sk.sitkNearestNeighbor] * 8)
one = reg.apply_transform(labels[0], fixed, tfm, 0, sk.sitkNearestNeighbor)  # single label: resample_nn_on_grid_dvf
assert np.array_equal(one.array, outs[1].array)
atlas = {str(a): {"DIR": {"S": Image(np.roll(labels[0].array, a - 1, axis=2), sp), "Weight Map": fusion.compute_weight_map(fixed, moving, "local", fusion.DEFAULT_VOTE_PARAMS)}}
         for a in range(3)}
prob = fusion.combine_labels(atlas, "S")["S"]
mask = fusion.process_probability_image(prob, 0.5)
st = fusion.combine_labels_staple({a: {"S": atlas[a]["DIR"]["S"]} for a in atlas})["S"]
with torch.cuda.stream(eng.stream):
    packed = torch.zeros(labels[0].array.shape, dtype=torch.uint8, device=eng.device)
for a in range(3):
    eng.pack_label(eng.to_device(atlas[str(a)]["DIR"]["S"]), a, packed, False)
w = eng.to_host(eng.staple_packed(packed, 0b111, eng.to_device(fixed)))
assert np.array_equal(w.array, st.array)
_, ltfm = linear.linear_registration(fixed, moving, reg_method="rigid", shrink_factors=[4, 2], smooth_sigmas=[2, 0], sampling_rate=0.5, number_of_iterations=6,
                                     metric="mattes_mi", default_value=-1000)
eng.synchronize()
print("SANITIZE RUN OK", float(np.abs(dvf.array).max()), int(mask.array.sum()))
