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


def test_run_segmentation_with_linear_prealignment(engine):
    """Atlases in their own space (shifted / rotated copies): linear_registration -> label propagation -> Demons ->
    fusion -> process_probability_image (multiatlas/run.py:261-404).  Functional bar of the reference's own test
    (platipy/imaging/tests/test_cardiac.py:231,237): Dice > 0.9 against the structure the atlases were made from."""
    from platipy_b200 import registration as reg
    from platipy_b200 import linear

    size, sp = (96, 64, 48), (1.0, 1.0, 1.5)
    target, _ = synth_pair(size, seed=3, spacing=sp, peak_mm=2.0)
    truth = [Image(l, sp) for l in synth_labels(size, 2, seed=520)]
    c = linear.image_center(target)
    atlas_set = {}
    for a, (ang, shift) in enumerate([(0.04, (3.0, -2.0, 1.5)), (-0.03, (-2.5, 2.0, -1.0)), (0.02, (1.0, 3.0, 2.0))]):
        ca, sa = np.cos(ang), np.sin(ang)
        t = sk.AffineTransform([[ca, -sa, 0], [sa, ca, 0], [0, 0, 1]], shift, c)
        inv = sk.AffineTransform(np.linalg.inv(t.matrix), -np.linalg.inv(t.matrix) @ t.offset, (0, 0, 0))
        _, ct = synth_pair(size, seed=3, spacing=sp, peak_mm=2.0, moving_seed=200 + a)  # target anatomy, deformed a little
        entry = {"CT Image": reg.apply_transform(ct, target, inv, -1000, sk.sitkLinear)}
        for k, lab in enumerate(truth):
            entry[f"S{k}"] = reg.apply_transform(lab, target, inv, 0, sk.sitkNearestNeighbor)
        atlas_set[f"{a:03d}"] = entry
    settings = {
        "linear_registration_settings": {"reg_method": "rigid", "shrink_factors": [4, 2], "smooth_sigmas": [2, 1], "sampling_rate": 0.5,
                                         "default_value": -1000, "number_of_iterations": 50, "metric": "mean_squares",
                                         "optimiser": "gradient_descent_line_search", "verbose": False},
        "deformable_registration_settings": {"isotropic_resample": False, "resolution_staging": [2, 1], "iteration_staging": [20, 10],
                                             "ncores": 8, "default_value": -1000, "verbose": False},
        "label_fusion_settings": {"vote_type": "unweighted", "vote_params": None, "optimal_threshold": {}, "fusion": "vote"},
    }
    results, probs = multiatlas.run_segmentation(target, atlas_set, settings)
    for k, lab in enumerate(truth):
        a, b = results[f"S{k}"].array > 0, lab.array > 0
        dice = 2.0 * (a & b).sum() / max(a.sum() + b.sum(), 1)
        print(f"S{k}: Dice {dice:.3f}")
        assert dice > 0.9, (k, dice)
    # the whole reference flow: auto-crop (run.py:200-246), paste back (run.py:387-404), post-processing (run.py:409-437)
    settings["auto_crop_target_image_settings"] = {"expansion_mm": [4, 4, 6]}
    settings["postprocessing_settings"] = {"run_postprocessing": True, "binaryfillhole_mm": 2, "structures_for_binaryfillhole": ["S0"],
                                           "structures_for_overlap_correction": ["S0", "S1"]}
    results2, probs2 = multiatlas.run_segmentation(target, atlas_set, settings)
    for k, lab in enumerate(truth):
        assert results2[f"S{k}"].GetSize() == target.GetSize() and probs2[f"S{k}"].GetSize() == target.GetSize()
        assert results2[f"S{k}"].GetPixelID() == sk.sitkUInt8
        a, b = results2[f"S{k}"].array > 0, lab.array > 0
        dice = 2.0 * (a & b).sum() / max(a.sum() + b.sum(), 1)
        print(f"S{k} (auto-crop + post-processing): Dice {dice:.3f}")
        assert dice > 0.88, (k, dice)
    assert not np.any((results2["S0"].array > 0) & (results2["S1"].array > 0))


def test_run_segmentation_from_disk_like_the_reference(engine, tmp_path):
    """``run_segmentation(img, settings)`` with an ``atlas_settings`` block (run.py:106-190): atlases are NIfTI files named by
    the format strings, optionally cropped to their structures; the result equals the in-memory call on the same data."""
    target, atlas_set = make_case(n_atlas=2)
    names = sorted(k for k in atlas_set["000"] if k != "CT Image")
    for a, entry in atlas_set.items():
        (tmp_path / f"Case_{a}" / "Images").mkdir(parents=True)
        (tmp_path / f"Case_{a}" / "Structures").mkdir(parents=True)
        sk.WriteImage(entry["CT Image"], str(tmp_path / f"Case_{a}" / "Images" / f"Case_{a}_CROP.nii.gz"))
        for n in names:
            sk.WriteImage(entry[n], str(tmp_path / f"Case_{a}" / "Structures" / f"Case_{a}_{n}_CROP.nii.gz"))
    settings = {
        "atlas_settings": {"atlas_id_list": sorted(atlas_set), "atlas_structure_list": names, "atlas_path": str(tmp_path),
                           "atlas_image_format": "Case_{0}/Images/Case_{0}_CROP.nii.gz",
                           "atlas_label_format": "Case_{0}/Structures/Case_{0}_{1}_CROP.nii.gz",
                           "crop_atlas_to_structures": False, "crop_atlas_expansion_mm": (20, 20, 40)},
        "deformable_registration_settings": {"isotropic_resample": False, "resolution_staging": [2, 1], "iteration_staging": [8, 4], "ncores": 8,
                                             "default_value": -1000, "verbose": False},
        "label_fusion_settings": {"vote_type": "unweighted", "vote_params": None, "optimal_threshold": {}},
    }
    res_disk, prob_disk = multiatlas.run_segmentation(target, settings)            # the reference's call form
    res_mem, prob_mem = multiatlas.run_segmentation(target, atlas_set, settings)
    for n in names:
        assert np.array_equal(res_disk[n].array, res_mem[n].array) and np.array_equal(prob_disk[n].array, prob_mem[n].array)
    # crop_atlas_to_structures: the loader crops image and structures to the structures' bounding box + expansion
    settings["atlas_settings"].update(crop_atlas_to_structures=True, crop_atlas_expansion_mm=(6, 6, 6))
    loaded = multiatlas.load_atlas_set(settings)
    exp_size, exp_index = ref.label_to_roi([atlas_set["000"][n] for n in names], [6, 6, 6])
    assert loaded["000"]["CT Image"].GetSize() == tuple(exp_size)
    assert np.array_equal(loaded["000"][names[0]].array, ref.crop_to_roi(atlas_set["000"][names[0]], exp_size, exp_index).array)
