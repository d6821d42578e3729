"""CPU checks of the oracle's binary post-processing (fusion.py:295-328) against scipy.ndimage and hand-made cases."""
import numpy as np
import scipy.ndimage as ndi

from oracle import itk_oracle as orc
from oracle import platipy_ref as ref
from platipy_b200.sitk_compat import Image


def _blobs(shape, seed, thr=0.55):
    rng = np.random.default_rng(seed)
    v = ndi.gaussian_filter(rng.standard_normal(shape), 2.0)
    v = (v - v.min()) / (v.max() - v.min())
    return (v > thr).astype(np.uint8)


def test_fillhole_matches_scipy_face_connectivity():
    for seed, shape in [(1, (20, 33, 47)), (2, (9, 64, 70)), (3, (1, 40, 40)), (4, (30, 30, 30))]:
        m = _blobs(shape, seed, 0.5)
        got = orc.binary_fillhole(m)
        exp = ndi.binary_fill_holes(m).astype(np.uint8)  # default structure: face connectivity of the background
        assert np.array_equal(got, exp)
    # a hole that touches the background only through an edge (diagonal) stays a hole with face connectivity
    m = np.zeros((3, 5, 5), np.uint8)
    m[0, 1:4, 1:4] = 1
    m[2, 1:4, 1:4] = 1
    m[1, 1:4, 1:4] = 1
    m[1, 2, 2] = 0
    m[1, 1, 1] = 0  # corner of the ring removed: the centre now touches the outside diagonally only
    got = orc.binary_fillhole(m)
    assert got[1, 2, 2] == 1 and got[1, 1, 1] == 0
    # a single slice has no holes in 3-D: every voxel touches the (padded) border through z
    m = np.zeros((1, 5, 5), np.uint8)
    m[0, 1:4, 1:4] = 1
    m[0, 2, 2] = 0
    assert orc.binary_fillhole(m)[0, 2, 2] == 0


def test_largest_component_matches_scipy_and_breaks_ties_in_raster_order():
    for seed, shape in [(5, (20, 33, 47)), (6, (12, 64, 31))]:
        m = _blobs(shape, seed, 0.6)
        got, ncomp, nvox, labels = orc.largest_component(m, want_labels=True)
        lab, n = ndi.label(m)  # face connectivity, labels in raster order of the first voxel
        assert ncomp == n
        assert np.array_equal(labels, lab)
        counts = np.bincount(lab.ravel())[1:]
        k = int(np.argmax(counts)) + 1
        assert nvox == counts[k - 1]
        assert np.array_equal(got, (lab == k).astype(np.uint8))
    m = np.zeros((2, 4, 9), np.uint8)
    m[0, 0, 0:3] = 1
    m[1, 3, 5:8] = 1  # same size, later in raster order
    got, ncomp, nvox = orc.largest_component(m)
    assert ncomp == 2 and nvox == 3 and got[0, 0, 1] == 1 and got[1, 3, 6] == 0
    empty, ncomp, nvox = orc.largest_component(np.zeros((3, 3, 3), np.uint8))
    assert ncomp == 0 and nvox == 0 and empty.sum() == 0


def test_process_probability_image_steps():
    p = np.zeros((6, 12, 12), np.float32)
    p[1:5, 2:9, 2:9] = 0.8     # big object ...
    p[2:4, 4:7, 4:7] = 0.1     # ... with a hole
    p[1:3, 10:12, 10:12] = 0.4  # small object, touches the border
    out = ref.process_probability_image(Image(p), 0.45)
    assert out.array.dtype == np.uint8
    exp = np.zeros_like(p, dtype=np.uint8)
    exp[1:5, 2:9, 2:9] = 1     # hole filled, small object (0.4 / 0.8 = 0.5 >= 0.45) removed as the smaller component
    assert np.array_equal(out.array, exp)
    assert ref.process_probability_image(np.zeros((3, 4, 5), np.float32)).array.sum() == 0


def test_box_mean_matches_scipy_cropped_window():
    # sitk.BoxMean (fusion.py:190): window cropped at the border, divided by the pixels inside
    a = np.random.default_rng(0).random((9, 12, 14)).astype(np.float32)
    got = ref.box_mean_f32(a, (2, 3, 1))  # radius (x, y, z)
    size = (3, 7, 5)                      # scipy window (z, y, x)
    s = ndi.uniform_filter(a.astype(np.float64), size=size, mode="constant", cval=0.0)
    c = ndi.uniform_filter(np.ones(a.shape), size=size, mode="constant", cval=0.0)
    assert np.allclose(got, (s / c).astype(np.float32), rtol=3e-7, atol=0)
    assert np.array_equal(ref.box_mean_f32(a, (0, 0, 0)), a)


def test_bspline3_matches_scipy_mirror_spline():
    """itk::BSplineInterpolateImageFunction (order 3) restatement vs scipy's cubic spline with mirror boundary
    (same pole, same mirror extension; ITK truncates the causal initialisation at 1e-10)."""
    from platipy_b200 import sitk_compat as sk

    rng = np.random.default_rng(0)
    a = rng.standard_normal((14, 30, 40)).astype(np.float32)
    g = orc.make_geom((40, 30, 14), (1, 1, 1), (0, 0, 0), (1, 0, 0, 0, 1, 0, 0, 0, 1))
    c = orc.bspline3_coefficients(a, g)
    assert np.allclose(c, ndi.spline_filter(a.astype(np.float64), order=3, mode="mirror"), rtol=0, atol=1e-8)
    go = orc.make_geom((37, 29, 13), (1.03, 0.97, 1.01), (0.4, 0.3, 0.2), (1, 0, 0, 0, 1, 0, 0, 0, 1))
    out = orc.resample_scalar(a, g, go, (), sk.sitkBSpline, 0.0)
    zz, yy, xx = np.meshgrid(np.arange(13) * 1.01 + 0.2, np.arange(29) * 0.97 + 0.3, np.arange(37) * 1.03 + 0.4, indexing="ij")
    exp = ndi.map_coordinates(a.astype(np.float64), [zz, yy, xx], order=3, mode="mirror")
    assert np.allclose(out, exp, rtol=0, atol=1e-6)
    short = rng.standard_normal((3, 5, 9))
    gs = orc.make_geom((9, 5, 3), (1, 1, 1), (0, 0, 0), (1, 0, 0, 0, 1, 0, 0, 0, 1))
    assert np.allclose(orc.bspline3_coefficients(short, gs), ndi.spline_filter(short, order=3, mode="mirror"), rtol=0, atol=1e-12)
    # interpolating splines reproduce the samples on the grid
    same = orc.resample_scalar(a, g, g, (), sk.sitkBSpline, 0.0)
    assert np.allclose(same, a, rtol=0, atol=2e-6)
