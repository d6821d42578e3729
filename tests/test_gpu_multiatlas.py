"""Multi-atlas pipeline (Demons -> batched label propagation -> weight map -> fusion) on one GPU against the
oracle pipeline run serially on the CPU; the sharded (N > 1) arithmetic is identical by construction except for
the association of float32 sums (unweighted votes are exact) -- see tests/dist_check.py for the 2-GPU run."""
import numpy as np
import pytest

from oracle import platipy_ref as ref
from platipy_b200 import multiatlas
from platipy_b200 import sitk_compat as sk
from platipy_b200.sitk_compat import Image
from platipy_b200.synth import synth_labels, synth_pair

pytestmark = pytest.mark.gpu


def make_case(size=(48, 40, 28), n_atlas=3, n_struct=2):
    sp = (1.0, 1.0, 1.5)
    target, _ = synth_pair(size, seed=0, spacing=sp, peak_mm=3.0)
    base = synth_labels(size, n_struct, seed=500)
    atlas_set = {}
    for a in range(n_atlas):
        _, ct = synth_pair(size, seed=0, spacing=sp, peak_mm=3.0, moving_seed=100 + a)
        entry = {"CT Image": ct}
        for s in range(n_struct):
            entry[f"S{s}"] = Image(np.roll(base[s], a - 1, axis=2), sp)
        atlas_set[f"{a:03d}"] = entry
    return target, atlas_set


def oracle_pipeline(target, atlas_set, settings):
    dset = settings["deformable_registration_settings"]
    kw = {k: v for k, v in dset.items() if k not in ("ncores", "verbose")}
    out = {}
    for a in sorted(atlas_set):
        _, tfm, _ = ref.fast_symmetric_forces_demons_registration(target, atlas_set[a]["CT Image"], **kw)
        d = {"CT Image": ref.apply_transform(atlas_set[a]["CT Image"], target, tfm, -1000, sk.sitkLinear)}
        for k, v in atlas_set[a].items():
            if k != "CT Image":
                d[k] = ref.apply_transform(v, target, tfm, 0, sk.sitkNearestNeighbor)
        d["Weight Map"] = ref.compute_weight_map(target, d["CT Image"], settings["label_fusion_settings"]["vote_type"], settings["label_fusion_settings"]["vote_params"])
        out[a] = {"DIR": d}
    return out


@pytest.mark.parametrize("fusion_mode", ["vote", "staple"])
def test_run_segmentation_matches_oracle_pipeline(engine, fusion_mode):
    target, atlas_set = make_case()
    settings = {
        "deformable_registration_settings": {"isotropic_resample": False, "resolution_staging": [2, 1], "iteration_staging": [8, 4], "ncores": 8,
                                             "default_value": -1000, "verbose": False},
        "label_fusion_settings": {"vote_type": "unweighted", "vote_params": None, "optimal_threshold": {"S0": 0.5}, "fusion": fusion_mode},
    }
    results, probs = multiatlas.run_segmentation(target, atlas_set, settings)
    dirs = oracle_pipeline(target, atlas_set, settings)
    if fusion_mode == "vote":
        exp = ref.combine_labels(dirs, ["S0", "S1"])
        for s in exp:
            assert np.array_equal(probs[s].array, exp[s].array), s
    else:
        exp = ref.combine_labels_staple({a: {k: v for k, v in dirs[a]["DIR"].items() if k.startswith("S")} for a in dirs})
        for s in exp:
            assert np.allclose(probs[s].array, exp[s].array, rtol=1e-9, atol=1e-12), s
    for s in results:
        assert results[s].GetPixelID() == sk.sitkUInt8 and results[s].GetSize() == target.GetSize()
        # run.py:373-384: process_probability_image(probability_map, optimal_threshold) -- on the GPU's own fused map
        # (bit-exact integer post-processing; the STAPLE map itself carries the 1e-9 reduction-order tolerance)
        thr = settings["label_fusion_settings"]["optimal_threshold"].get(s, 0.5)
        assert np.array_equal(results[s].array, ref.process_probability_image(probs[s], thr).array), s
