"""Committed fixtures (tests/golden/, oracle-generated -- see make_golden.py): the oracle must keep reproducing
them (CPU), and the CUDA path must match them (GPU)."""
import os

import numpy as np
import pytest

from platipy_b200 import sitk_compat as sk
from platipy_b200.sitk_compat import Image

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "demons_small.npz"))


def _inputs():
    sp, og = tuple(G["spacing"]), tuple(G["origin"])
    kw = dict(resolution_staging=[int(v) for v in G["resolution_staging"]], iteration_staging=[int(v) for v in G["iteration_staging"]])
    return Image(G["fixed"], sp, og), Image(G["moving"], sp, og), Image(G["label"], sp, og), kw


def test_oracle_reproduces_golden():
    from oracle import platipy_ref as ref

    F, M, L, kw = _inputs()
    reg, tfm, dvf = ref.fast_symmetric_forces_demons_registration(F, M, **kw)
    assert np.array_equal(dvf.array, G["dvf"]) and np.array_equal(reg.array, G["registered"])
    assert np.array_equal(ref.apply_transform(L, F, tfm, 0, sk.sitkNearestNeighbor).array, G["warped_label"])


@pytest.mark.gpu
def test_gpu_matches_golden(engine):
    from platipy_b200 import registration as reg

    F, M, L, kw = _inputs()
    img, tfm, dvf = reg.fast_symmetric_forces_demons_registration(F, M, **kw)
    assert np.abs(dvf.array - G["dvf"]).max() <= 1e-4  # mm
    assert np.abs(img.array - G["registered"]).max() <= 1e-5 * np.abs(G["registered"]).max()
    assert np.array_equal(reg.apply_transform(L, F, tfm, 0, sk.sitkNearestNeighbor).array, G["warped_label"])


# ---- second fixture: the rows around the Demons loop --------------------------------------------------------------
R = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rows_small.npz"))
BLOCK_PARAMS = {"factor": 1e12, "gain": 6, "blockSize": (2, 2, 1), "normalise": True}


def _row_inputs():
    sp, og = tuple(G["spacing"]), tuple(G["origin"])
    F, M = Image(G["fixed"], sp, og), Image(G["moving"], sp, og)
    tfm = sk.DisplacementFieldTransform(Image(G["dvf"], sp, og, is_vector=True))
    prob = Image(R["prob"], sp, og)
    labs = {"A": Image((R["prob"] > 0.5).astype(np.uint8), sp, og), "B": Image((np.roll(R["prob"], 3, axis=2) > 0.55).astype(np.uint8), sp, og)}
    return F, M, tfm, prob, labs, og


def test_oracle_reproduces_row_golden():
    from oracle import platipy_ref as ref

    F, M, tfm, prob, labs, og = _row_inputs()
    assert np.array_equal(ref.apply_transform(M, F, tfm, -1000, sk.sitkBSpline).array, R["bspline"])
    assert np.array_equal(ref.process_probability_image(prob, 0.45).array, R["mask"])
    assert np.array_equal(ref.compute_weight_map(F, M, "block", BLOCK_PARAMS).array, R["block"])
    fx = ref.correct_volume_overlap(labs)
    assert np.array_equal(fx["A"].array, R["overlap_a"]) and np.array_equal(fx["B"].array, R["overlap_b"])
    assert ref.label_to_roi(labs["A"], [1, 1, 2.5], return_as_list=True) == [int(v) for v in R["roi"]]
    acc = ref.linreg_meansq(F, M, R["lin_matrix"], R["lin_offset"], np.eye(3), np.array(og), stride=3)
    assert np.allclose(acc, R["lin_acc"], rtol=1e-12, atol=0)


@pytest.mark.gpu
def test_gpu_matches_row_golden(engine):
    from platipy_b200 import fusion, label_utils as lu
    from platipy_b200 import registration as reg

    F, M, tfm, prob, labs, og = _row_inputs()
    assert np.array_equal(reg.apply_transform(M, F, tfm, -1000, sk.sitkBSpline).array, R["bspline"])
    assert np.array_equal(fusion.process_probability_image(prob, 0.45).array, R["mask"])
    assert np.allclose(fusion.compute_weight_map(F, M, "block", BLOCK_PARAMS).array, R["block"], rtol=1e-5, atol=0)
    fx = lu.correct_volume_overlap(labs)
    assert np.array_equal(fx["A"].array, R["overlap_a"]) and np.array_equal(fx["B"].array, R["overlap_b"])
    assert lu.label_to_roi(labs["A"], [1, 1, 2.5], return_as_list=True) == [int(v) for v in R["roi"]]
    acc = engine.linreg_meansq(engine.to_device(F), engine.to_device(M), R["lin_matrix"], R["lin_offset"], np.eye(3), np.array(og), None, None, 3)
    assert acc[1] == R["lin_acc"][1] and np.allclose(acc, R["lin_acc"], rtol=1e-9, atol=1e-9 * np.abs(R["lin_acc"]).max())


# ---- third fixture: distance maps, contours, morphology, generators, patch correlation, metric sums, surface metrics ----------------
def test_oracle_reproduces_rows3_golden():
    """(The GPU counterpart lives in tests/test_gpu_zzz_session3.py with the other tests of these rows.)"""
    import importlib.util

    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(here, "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    now = mg.main_rows3(write=False)
    ref3 = np.load(os.path.join(here, "golden", "rows3_small.npz"))
    assert sorted(now) == sorted(ref3.files)
    for k in ref3.files:
        a, b = np.asarray(now[k]), ref3[k]
        if a.dtype.kind in "fc" and k in ("patch_weight", "corr_sums", "mattes_hist", "surface_metric_values"):  # numpy / scipy reductions
            assert np.allclose(a, b, rtol=1e-10, atol=1e-12), k
        else:
            assert np.array_equal(a, b), k
