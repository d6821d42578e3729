"""The reference's own known-answer tests for the surface metrics (platipy/imaging/tests/test_metrics.py:6-67), restated against
the oracle.  The eleven golden numbers below were produced by the real SimpleITK (they are asserted in the reference's test-suite);
reproducing them PINS the oracle's restatements of sitk.SignedMaurerDistanceMap, BinaryContour / LabelContour,
HausdorffDistanceImageFilter and LabelIntensityStatisticsImageFilter against ITK itself -- the single place where this repository
has golden vectors of the reference to anchor on.  The GPU path is then held to the oracle bit for bit
(tests/test_gpu_zz_generation.py) and to the same golden numbers (tests/test_gpu_zzz_session3.py)."""
import numpy as np

from oracle import comparison_ref as cref
from platipy_b200.sitk_compat import Image


def cube(lo, hi):
    """sitk.Image(100, 100, 100, sitkUInt8) with spacing (1, 1, 2) and label[lo:hi, lo:hi, lo:hi] = 1 (test_metrics.py:8-10)."""
    arr = np.zeros((100, 100, 100), np.uint8)
    arr[lo:hi, lo:hi, lo:hi] = 1  # the same range on every axis: (x, y, z) and [z, y, x] indexing agree
    return Image(arr, (1.0, 1.0, 2.0))


def test_surface_dsc():
    # test_metrics.py:6-38
    label_a = cube(30, 70)
    assert cref.compute_surface_dsc(label_a, cube(30, 71)) == 1.0
    assert np.allclose(cref.compute_surface_dsc(label_a, cube(35, 71)), 0.5158373786407767)
    assert np.allclose(cref.compute_surface_dsc(label_a, cube(35, 72)), 0.39725541227966404)
    assert np.allclose(cref.compute_surface_dsc(label_a, cube(35, 75)), 0.1258764241893076)


def test_surface_metrics():
    # test_metrics.py:40-67
    label_a = cube(30, 70)
    metrics = cref.compute_surface_metrics(label_a, cube(30, 71))
    assert np.allclose(metrics["hausdorffDistance"], 2.449489742783178)
    assert np.allclose(metrics["meanSurfaceDistance"], 0.6649174304423457)
    assert np.allclose(metrics["medianSurfaceDistance"], 0.574099183082580)
    assert np.allclose(metrics["maximumSurfaceDistance"], 2.4494898319244385)
    assert np.allclose(metrics["sigmaSurfaceDistance"], 101.78549149738755)
    assert np.allclose(metrics["surfaceDSC"], 1.0)
    metrics = cref.compute_surface_metrics(label_a, cube(35, 71))
    assert np.allclose(metrics["hausdorffDistance"], 12.24744871391589)
    assert np.allclose(metrics["meanSurfaceDistance"], 3.842314521867095)
    assert np.allclose(metrics["medianSurfaceDistance"], 3.5163573920726776)
    assert np.allclose(metrics["maximumSurfaceDistance"], 12.24744871391589)
    assert np.allclose(metrics["sigmaSurfaceDistance"], 392.57229390698296)
    assert np.allclose(metrics["surfaceDSC"], 0.5158373786407767)


def test_golden_values_to_the_last_digit():
    """np.allclose (the reference's own criterion) allows 1e-5; the restatement is much closer than that: the float64 means of the
    float32 distances are identical, i.e. the distance maps agree voxel for voxel on both contours."""
    label_a = cube(30, 70)
    m1, m2 = cref.compute_surface_metrics(label_a, cube(30, 71)), cref.compute_surface_metrics(label_a, cube(35, 71))
    assert m1["meanSurfaceDistance"] == 0.6649174304423457 and m2["meanSurfaceDistance"] == 3.842314521867095
    assert m1["maximumSurfaceDistance"] == 2.4494898319244385  # sqrt(6) in float32 (LabelIntensityStatistics on the Float32 map)
    assert m1["hausdorffDistance"] == 2.449489742783178 and m2["hausdorffDistance"] == 12.24744871391589  # sqrt(6), sqrt(150) in double
    assert abs(m1["sigmaSurfaceDistance"] - 101.78549149738755) < 1e-9 and abs(m2["sigmaSurfaceDistance"] - 392.57229390698296) < 1e-9
    assert cref.compute_surface_dsc(label_a, cube(35, 72)) == 0.39725541227966404
    # the median pins the histogram rule: bin width = global maximum of |distance| / 128, the value is a bin centre
    assert abs(m1["medianSurfaceDistance"] - 0.574099183082580) < 1e-12


def test_other_comparison_metrics_are_consistent():
    a, b = cube(30, 70), cube(35, 71)
    vm = cref.compute_volume_metrics(a, b)
    inter = 35 ** 3
    assert np.isclose(vm["DSC"], 2 * inter / (40 ** 3 + 36 ** 3)) and np.isclose(vm["volumeOverlap"], inter * 2 / 1000.0)
    assert cref.compute_metric_dsc(a, b) == vm["DSC"] == cref.compute_metric_dsc(a, b, auto_crop=False)
    assert np.isclose(cref.compute_volume(a), 40 ** 3 * 2 / 1000.0)
    assert cref.compute_metric_hd(a, b, auto_crop=False) == cref.hausdorff_distance(a, b)
    assert np.isclose(cref.compute_metric_masd(a, b, auto_crop=False), 3.842314521867095)
    assert np.isnan(cref.compute_metric_hd(a, Image(np.zeros_like(a.array), a.GetSpacing())))
    assert 0 < cref.compute_metric_sensitivity(a, b) < 1 and 0 < cref.compute_metric_specificity(a, b) <= 1


def test_added_path_length_restatement_and_emulated_slicewise_contour(emu):
    import ctypes as C

    import scipy.ndimage as ndi
    from oracle import itk_oracle as orc

    a, b = cube(30, 70), cube(35, 71)
    # identical labels add nothing; a shifted cube adds the part of the reference outline farther than the threshold
    assert sum(cref.compute_apl(a, a)) == 0 and len(cref.compute_apl(a, a)) == 40
    apl = cref.compute_apl(a, b, 3)
    assert len(apl) == 41 and sum(apl) > 0 and cref.compute_metric_total_apl(a, b, 3) == sum(apl) * 1.0
    assert sum(cref.compute_apl(a, b, 0)) >= sum(apl)
    # slices 30..34 hold the reference alone: its whole outline (4 * 40 - 4 voxels) is added path
    assert all(int(v) == 156 for v in apl[:5])
    # the slice-wise contour kernel (host emulation) against the per-slice restatement
    r = np.random.default_rng(3)
    lab = (ndi.gaussian_filter(r.standard_normal((9, 30, 34)), 2.0) > 0.02).astype(np.uint8)
    out = np.empty_like(lab)
    emu.emu_label_contour_slicewise(lab.ctypes.data_as(C.c_void_p), 34, 30, 9, out.ctypes.data_as(C.c_void_p), C.c_uint(3), C.c_uint(64))
    exp = np.stack([orc.label_contour(lab[i:i + 1], False)[0] for i in range(9)])
    assert np.array_equal(out, exp) and not np.array_equal(out, orc.label_contour(lab, False))
