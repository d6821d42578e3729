"""GPU parity of the label utilities around the fusion step (utils/crop.py:24-99, label/utils.py:23-58, sitk.Paste and
sitk.BinaryMorphologicalClosing as used in multiatlas/run.py:387-437) against the oracle's numpy / scipy restatement --
integer work, bit-exact."""
import numpy as np
import pytest
import scipy.ndimage as ndi

from oracle import platipy_ref as ref
from platipy_b200 import label_utils as lu
from platipy_b200.sitk_compat import Image
from platipy_b200.synth import synth_labels

pytestmark = pytest.mark.gpu


def test_label_to_roi_crop_and_paste(engine):
    size, sp = (50, 44, 30), (1.0, 1.5, 2.5)
    labs = [Image(l, sp, (3.0, -2.0, 7.0)) for l in synth_labels(size, 3, seed=610)]
    for args in ((labs[0], [0, 0, 0]), (labs[1], [5, 5, 10]), (labs, [20, 20, 40]), (labs[2], [100, 100, 100])):
        assert lu.label_to_roi(args[0], args[1]) == ref.label_to_roi(args[0], args[1])
    assert lu.label_to_roi(labs[0], [4, 4, 4], return_as_list=True) == ref.label_to_roi(labs[0], [4, 4, 4], return_as_list=True)
    with pytest.raises(RuntimeError):
        lu.label_to_roi(Image(np.zeros((4, 4, 4), np.uint8)))
    rng = np.random.default_rng(0)
    for dtype in (np.float32, np.uint8, np.int16, np.float64):
        img = Image((rng.random((30, 44, 50)) * 100).astype(dtype), sp, (3.0, -2.0, 7.0))
        csize, cidx = ref.label_to_roi(labs, [6, 6, 6])
        got, exp = lu.crop_to_roi(img, csize, cidx), ref.crop_to_roi(img, csize, cidx)
        assert np.array_equal(got.array, exp.array) and got.array.dtype == dtype
        assert np.allclose(got.GetOrigin(), exp.GetOrigin()) and got.GetSpacing() == img.GetSpacing()
        template = Image(np.zeros_like(img.array), sp, (3.0, -2.0, 7.0))
        back = lu.paste(template, got, got.GetSize(), (0, 0, 0), cidx)
        assert np.array_equal(back.array, ref.paste(template, exp, exp.GetSize(), (0, 0, 0), cidx).array)
        assert np.array_equal(lu.crop_to_label_extent(img, labs[0], 3).array, ref.crop_to_roi(img, *ref.label_to_roi(labs[0], [3, 3, 3])).array)
    with pytest.raises(RuntimeError):
        lu.crop_to_roi(labs[0], (60, 10, 10), (0, 0, 0))


def test_correct_volume_overlap_bit_exact(engine):
    size = (40, 36, 24)
    base = synth_labels(size, 4, seed=620)
    d = {f"S{k}": Image(np.roll(l, k * 2, axis=2), (1.0, 1.0, 2.0)) for k, l in enumerate(base)}
    for largest in (True, False):
        got, exp = lu.correct_volume_overlap(d, largest), ref.correct_volume_overlap(d, largest)
        assert list(got) == list(exp)
        tot = np.zeros(base[0].shape, np.int32)
        for k in got:
            assert np.array_equal(got[k].array, exp[k].array), k
            tot += got[k].array
        assert tot.max() <= 1  # no overlap left


def test_binary_closing_matches_scipy(engine):
    rng = np.random.default_rng(4)
    v = ndi.gaussian_filter(rng.standard_normal((24, 40, 44)), 2.0)
    m = Image((v > 0.02).astype(np.uint8), (1.0, 1.0, 3.0))
    for radius in ((3, 3, 1), (1, 1, 1), (2, 0, 1), 2):
        r = [radius] * 3 if np.isscalar(radius) else list(radius)
        offs = lu.ball_offsets(r)
        st = np.zeros((2 * r[2] + 1, 2 * r[1] + 1, 2 * r[0] + 1), bool)
        st[offs[:, 2] + r[2], offs[:, 1] + r[1], offs[:, 0] + r[0]] = True
        assert st[r[2], r[1], r[0]] and np.array_equal(st, st[::-1, ::-1, ::-1])
        got = lu.binary_morphological_closing(m, radius)
        exp = ref.binary_morphological_closing(m, r, st)
        assert np.array_equal(got.array, exp.array), radius
        assert np.all(got.array >= m.array)  # closing is extensive
