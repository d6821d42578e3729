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
