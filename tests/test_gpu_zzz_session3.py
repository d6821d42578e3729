"""GPU tests of the rows written AFTER the round's last GPU minute (third session of round 1).  Everything here has been checked
on the CPU only: the kernels under the host emulation of tests/emu (bit for bit against the oracle), the host logic against the
oracle metrics, the oracle itself against the reference's golden values where those exist.  The device pieces they build on
(distance map, contours, masking, bounding box, crop, resampling) did pass on a B200 (test_gpu_zz_generation.py,
test_gpu_label_utils.py).  The file name sorts last so that, under ``pytest -x``, these run after every test that has already been
seen green on the GPU.

* label comparison metrics (platipy/imaging/label/comparison.py): the first two tests ARE the reference's own known-answer tests
  (platipy/imaging/tests/test_metrics.py:6-67) with the import changed -- the golden numbers come from the real SimpleITK; then
  parity against the oracle on irregular labels, including the auto-crop and the added path length;
* compute_weight_map(vote_type="patch_correlation"); linear_registration with metric correlation / mattes_mi, optimiser lbfgsb;
  alignment_registration(moments=True); get_bone_mask.
"""
import numpy as np
import pytest
import scipy.ndimage as ndi

from oracle import comparison_ref as cref
from platipy_b200 import comparison as cmp
from platipy_b200 import generation as gen
from platipy_b200.label_utils import ball_offsets
from platipy_b200.sitk_compat import Image

pytestmark = pytest.mark.gpu


def cube(lo, hi):
    arr = np.zeros((100, 100, 100), np.uint8)
    arr[lo:hi, lo:hi, lo:hi] = 1
    return Image(arr, (1.0, 1.0, 2.0))


def test_surface_dsc(engine):
    label_a = cube(30, 70)
    assert cmp.compute_surface_dsc(label_a, cube(30, 71)) == 1.0
    assert np.allclose(cmp.compute_surface_dsc(label_a, cube(35, 71)), 0.5158373786407767)
    assert np.allclose(cmp.compute_surface_dsc(label_a, cube(35, 72)), 0.39725541227966404)
    assert np.allclose(cmp.compute_surface_dsc(label_a, cube(35, 75)), 0.1258764241893076)


def test_surface_metrics(engine):
    label_a = cube(30, 70)
    metrics = cmp.compute_surface_metrics(label_a, cube(30, 71))
    assert np.allclose(metrics["hausdorffDistance"], 2.449489742783178)
    assert np.allclose(metrics["meanSurfaceDistance"], 0.6649174304423457)
    assert np.allclose(metrics["medianSurfaceDistance"], 0.574099183082580)
    assert np.allclose(metrics["maximumSurfaceDistance"], 2.4494898319244385)
    assert np.allclose(metrics["sigmaSurfaceDistance"], 101.78549149738755)
    assert np.allclose(metrics["surfaceDSC"], 1.0)
    metrics = cmp.compute_surface_metrics(label_a, cube(35, 71))
    assert np.allclose(metrics["hausdorffDistance"], 12.24744871391589)
    assert np.allclose(metrics["meanSurfaceDistance"], 3.842314521867095)
    assert np.allclose(metrics["medianSurfaceDistance"], 3.5163573920726776)
    assert np.allclose(metrics["maximumSurfaceDistance"], 12.24744871391589)
    assert np.allclose(metrics["sigmaSurfaceDistance"], 392.57229390698296)
    assert np.allclose(metrics["surfaceDSC"], 0.5158373786407767)


def _blobs(shape, seed, level=0.02, sigma=2.5):
    r = np.random.default_rng(seed)
    return (ndi.gaussian_filter(r.standard_normal(shape), sigma) > level).astype(np.uint8)


def _largest(mask):
    lab, n = ndi.label(mask)
    return (lab == (1 + np.argmax(ndi.sum(mask, lab, range(1, n + 1))))).astype(np.uint8)


def test_metrics_match_the_oracle_on_irregular_labels(engine):
    shape, sp = (40, 56, 60), (0.9, 1.1, 2.5)
    a = Image(_largest(_blobs(shape, 21, 0.03, 4.0)), sp, (5.0, -3.0, 10.0))
    b = Image(np.roll(a.array, (1, -2, 3), axis=(0, 1, 2)) | _largest(_blobs(shape, 22, 0.05, 3.0)) & a.array, sp, (5.0, -3.0, 10.0))
    assert 0 < b.array.sum() and (a.array != b.array).any()
    got, exp = cmp.compute_surface_metrics(a, b), cref.compute_surface_metrics(a, b)
    assert set(got) == set(exp)
    for k in exp:
        assert np.isclose(got[k], exp[k], rtol=1e-12, atol=0), (k, got[k], exp[k])
    for tau in (1.0, 3.0, 7.5):
        assert cmp.compute_surface_dsc(a, b, tau) == cref.compute_surface_dsc(a, b, tau)
    gv, ev = cmp.compute_volume_metrics(a, b), cref.compute_volume_metrics(a, b)
    for k in ev:
        assert np.isclose(gv[k], ev[k], rtol=1e-14), k
    assert np.isclose(cmp.compute_volume(a), cref.compute_volume(a))
    for crop in (True, False):
        assert cmp.compute_metric_dsc(a, b, crop) == cref.compute_metric_dsc(a, b, crop)
        assert np.isclose(cmp.compute_metric_specificity(a, b, crop), cref.compute_metric_specificity(a, b, crop), rtol=1e-14)
        assert np.isclose(cmp.compute_metric_sensitivity(a, b, crop), cref.compute_metric_sensitivity(a, b, crop), rtol=1e-14)
        assert np.isclose(cmp.compute_metric_masd(a, b, crop), cref.compute_metric_masd(a, b, crop), rtol=1e-12)
        assert np.isclose(cmp.compute_metric_hd(a, b, crop), cref.compute_metric_hd(a, b, crop), rtol=1e-7)
    empty = Image(np.zeros(shape, np.uint8), sp, (5.0, -3.0, 10.0))
    assert np.isnan(cmp.compute_metric_hd(a, empty)) and np.isnan(cmp.compute_metric_masd(empty, a))
    # device in, same numbers
    da, db = engine.to_device(a), engine.to_device(b)
    assert cmp.compute_metric_dsc(da, db) == cref.compute_metric_dsc(a, b)
    assert np.isclose(cmp.compute_metric_hd(da, db), cref.compute_metric_hd(a, b), rtol=1e-7)
    for thr in (3, 0, 1.5):
        got_apl, exp_apl = cmp.compute_apl(a, b, thr), cref.compute_apl(a, b, thr)
        assert [int(v) for v in got_apl] == [int(v) for v in exp_apl] and len(got_apl) > 5
        assert np.isclose(cmp.compute_metric_total_apl(a, b, thr), cref.compute_metric_total_apl(a, b, thr))
        assert np.isclose(cmp.compute_metric_mean_apl(a, b, thr), cref.compute_metric_mean_apl(a, b, thr))
    assert sum(cmp.compute_apl(a, a, 3)) == 0


def test_patch_correlation_weight_map(engine):
    """vote_type "patch_correlation" (fusion.py:82-146): one kernel instead of a Python loop of scipy.stats.pearsonr calls.
    Float64 sums in window order vs numpy's pairwise / BLAS order: the Float32 weight map agrees to the north star's 1e-5."""
    from oracle import platipy_ref as ref
    from platipy_b200 import fusion
    from platipy_b200 import sitk_compat as sk
    from platipy_b200.synth import synth_pair

    t, m = synth_pair((40, 36, 24), seed=5, spacing=(1.0, 1.0, 2.0))
    for fn in (lambda x: x + 1, abs, lambda x: 0.5 * x + 1.5):
        vp = dict(patch_window_mm=12, resampled_voxel_size_mm=3, correlation_function=fn)
        got, exp = fusion.compute_weight_map(t, m, "patch_correlation", vp), ref.compute_weight_map(t, m, "patch_correlation", vp)
        assert got.GetPixelID() == sk.sitkFloat32 and got.array.shape == t.array.shape
        assert np.allclose(got.array, exp.array, rtol=1e-5, atol=1e-6)
    # a correlation function written against the host image API gets a host image
    vp["correlation_function"] = lambda x: Image(np.abs(x.array), x.GetSpacing(), x.GetOrigin(), x.GetDirection())
    got2 = fusion.compute_weight_map(t, m, "patch_correlation", vp)
    vp["correlation_function"] = abs
    assert np.array_equal(got2.array, fusion.compute_weight_map(t, m, "patch_correlation", vp).array)
    # device in -> device out, and the default parameters carry the reference's keys
    dw = fusion.compute_weight_map(engine.to_device(t), engine.to_device(m), "patch_correlation",
                                   dict(fusion.DEFAULT_VOTE_PARAMS, patch_window_mm=12, resampled_voxel_size_mm=3))
    vp["correlation_function"] = lambda x: x + 1
    assert np.array_equal(engine.to_host(dw).array, fusion.compute_weight_map(t, m, "patch_correlation", vp).array)


def test_linear_registration_correlation_metric(engine):
    """metric="correlation" (linear.py:141-146): the 42 sums of the kernel against the numpy restatement at identical poses, and a
    registration between images with different intensity scales (functional parity, like the mean-squares tests)."""
    from oracle import platipy_ref as ref
    from platipy_b200 import linear
    from platipy_b200 import sitk_compat as sk

    def blob(size, center, spacing=(1.0, 1.0, 1.0), sig=(7.0, 5.0, 4.0), origin=(0.0, 0.0, 0.0), direction=(1, 0, 0, 0, 1, 0, 0, 0, 1)):
        nx, ny, nz = size
        z, y, x = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
        p = [x * spacing[0], y * spacing[1], z * spacing[2]]
        v = sum(((pi - ci) / si) ** 2 for pi, ci, si in zip(p, center, sig))
        return Image((1000.0 * np.exp(-0.5 * v)).astype(np.float32), spacing, origin, direction)

    rng = np.random.default_rng(9)
    ang = 0.15
    rot = (np.cos(ang), -np.sin(ang), 0, np.sin(ang), np.cos(ang), 0, 0, 0, 1.0)
    f = blob((24, 20, 16), (12.0, 10.0, 8.0), (1.0, 1.2, 1.5))
    mv = blob((30, 26, 20), (15.0, 13.5, 12.0), (1.1, 1.0, 1.4), origin=(-3.0, 2.0, -1.0), direction=rot)
    fmask = Image((rng.random(f.array.shape) > 0.3).astype(np.uint8), f.GetSpacing())
    mmask = Image((rng.random(mv.array.shape) > 0.2).astype(np.uint8), mv.GetSpacing(), mv.GetOrigin(), mv.GetDirection())
    init = linear.centered_transform_initializer(f, mv)
    m = linear.make_model("affine")
    p = m.identity() + 0.02 * rng.standard_normal(m.n)
    A, b = init.matrix @ m.matrix(p), init.matrix @ m.offset(p) + init.offset
    df, dm, dfm, dmm = (engine.to_device(i) for i in (f, mv, fmask, mmask))
    for fm, mm, dfm_, dmm_, stride in ((None, None, None, None, 1), (fmask, None, dfm, None, 3), (fmask, mmask, dfm, dmm, 2)):
        got = engine.linreg_correlation(df, dm, A, b, init.matrix, m.center, dfm_, dmm_, stride)
        exp = ref.linreg_correlation(f, mv, A, b, init.matrix, m.center, fm, mm, stride)
        assert got.shape == (42,) and got[0] == exp[0]
        assert np.allclose(got, exp, rtol=1e-9, atol=1e-6 * np.abs(exp).max()), stride
        again = engine.linreg_correlation(df, dm, A, b, init.matrix, m.center, dfm_, dmm_, stride)
        assert np.array_equal(got, again)  # fixed-order sums: deterministic
    # moving = 2 * fixed + 100, shifted by (+2, -1.5, +1) mm
    fixed = blob((48, 40, 32), (24.0, 20.0, 16.0))
    shifted = blob((48, 40, 32), (26.0, 18.5, 17.0))
    moving = Image(shifted.array * 2.0 + 100.0, shifted.GetSpacing())
    for optimiser in ("gradient_descent", "gradient_descent_line_search"):
        registered, tfm = linear.linear_registration(fixed, moving, reg_method="translation", metric="correlation", optimiser=optimiser,
                                                     shrink_factors=[2, 1], smooth_sigmas=[1, 0], sampling_rate=0.5, number_of_iterations=60,
                                                     default_value=100)
        pt = np.array(tfm.flatten()[0].TransformPoint((24.0, 20.0, 16.0)))
        for t in tfm.flatten()[1:]:
            pt = np.array(t.TransformPoint(pt))
        assert np.allclose(pt, (26.0, 18.5, 17.0), atol=0.3), (optimiser, pt)
        r = np.corrcoef(registered.array.ravel(), fixed.array.ravel())[0, 1]
        assert r > 0.995 and min(linear.LAST_HISTORY[-1]) < -0.98
    with pytest.raises(NotImplementedError):
        linear.linear_registration(fixed, moving, metric="joint_hist_mi")


def test_get_bone_mask(engine):
    """generation/mask.py:21-47: BinaryThreshold then BinaryMorphologicalClosing with max_hole_size as the kernel radius."""
    from oracle import platipy_ref as ref

    rng = np.random.default_rng(2)
    ct = Image((ndi.gaussian_filter(rng.standard_normal((20, 40, 44)), 2.0) * 4000).astype(np.float32), (1.0, 1.0, 2.5))
    for hole in (2, (1, 2, 1)):
        got = gen.get_bone_mask(ct, 350, 3500, hole)
        thr = Image(((ct.array >= 350) & (ct.array <= 3500)).astype(np.uint8), ct.GetSpacing())
        r = [hole] * 3 if np.isscalar(hole) else list(hole)
        offs = ball_offsets(r)
        st = np.zeros((2 * r[2] + 1, 2 * r[1] + 1, 2 * r[0] + 1), bool)
        st[offs[:, 2] + r[2], offs[:, 1] + r[1], offs[:, 0] + r[0]] = True
        assert got.array.dtype == np.uint8 and np.array_equal(got.array, ref.binary_morphological_closing(thr, r, st).array)
        assert got.array.sum() >= thr.array.sum() > 0


def test_alignment_registration_with_moments_and_lbfgsb(engine):
    """alignment_registration(moments=True) (linear.py:23-47: CenteredTransformInitializer MOMENTS) and optimiser="lbfgsb"."""
    from oracle import platipy_ref as ref
    from platipy_b200 import linear

    def blob(size, center, spacing=(1.0, 1.0, 1.0), sig=(7.0, 5.0, 4.0), origin=(0.0, 0.0, 0.0)):
        nx, ny, nz = size
        z, y, x = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
        p = [x * spacing[0], y * spacing[1], z * spacing[2]]
        v = sum(((pi - ci) / si) ** 2 for pi, ci, si in zip(p, center, sig))
        return Image((1000.0 * np.exp(-0.5 * v)).astype(np.float32), spacing, origin)

    fixed = blob((48, 40, 32), (24.0, 20.0, 16.0))
    moving = blob((40, 44, 30), (17.0, 25.0, 13.0), spacing=(1.2, 1.0, 1.1), origin=(5.0, -8.0, 2.0))
    got = engine.image_moments(engine.to_device(moving))
    assert np.allclose(got, ref.image_moments(moving), rtol=1e-10)
    aligned, tfm = linear.alignment_registration(fixed, moving)  # moments=True is the reference's default
    # fixed blob centre -> moving blob centre (physical): (24, 20, 16) -> origin + (17, 25, 13)
    assert np.allclose(tfm.TransformPoint((24.0, 20.0, 16.0)), (22.0, 17.0, 15.0), atol=0.25)  # the moving blob is cut by its image border
    assert aligned.array.dtype == np.float32 and aligned.array.shape == fixed.array.shape
    assert np.corrcoef(aligned.array.ravel(), fixed.array.ravel())[0, 1] > 0.98
    with pytest.raises(RuntimeError):
        linear.alignment_registration(fixed, Image(np.zeros((8, 8, 8), np.float32)))
    shifted = blob((48, 40, 32), (26.0, 18.5, 17.0))
    for metric, mv in (("mean_squares", shifted), ("correlation", Image(shifted.array * 2.0 + 100.0, shifted.GetSpacing()))):
        _, t = linear.linear_registration(fixed, mv, reg_method="translation", metric=metric, optimiser="lbfgsb", shrink_factors=[2, 1],
                                          smooth_sigmas=[1, 0], sampling_rate=0.5, number_of_iterations=50)
        pt = np.array((24.0, 20.0, 16.0))
        for part in t.flatten():
            pt = np.array(part.TransformPoint(pt))
        assert np.allclose(pt, (26.0, 18.5, 17.0), atol=0.3), (metric, pt)


def test_linear_registration_mattes_mutual_information(engine):
    """metric="mattes_mi" (linear.py:145-146): histogram and derivative sums against the numpy restatement at identical poses, and
    a registration of an inverted-contrast pair, which only a mutual-information metric can align."""
    from oracle import platipy_ref as ref
    from platipy_b200 import linear

    def blob(size, center, spacing=(1.0, 1.0, 1.0), sig=(7.0, 5.0, 4.0), origin=(0.0, 0.0, 0.0), direction=(1, 0, 0, 0, 1, 0, 0, 0, 1)):
        nx, ny, nz = size
        z, y, x = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
        p = [x * spacing[0], y * spacing[1], z * spacing[2]]
        v = sum(((pi - ci) / si) ** 2 for pi, ci, si in zip(p, center, sig))
        return Image((1000.0 * np.exp(-0.5 * v)).astype(np.float32), spacing, origin, direction)

    rng = np.random.default_rng(9)
    ang = 0.15
    rot = (np.cos(ang), -np.sin(ang), 0, np.sin(ang), np.cos(ang), 0, 0, 0, 1.0)
    f = blob((24, 20, 16), (12.0, 10.0, 8.0), (1.0, 1.2, 1.5))
    big = blob((30, 26, 20), (15.0, 13.5, 12.0), (1.1, 1.0, 1.4), origin=(-3.0, 2.0, -1.0), direction=rot)
    mv = Image((900.0 - big.array).astype(np.float32), big.GetSpacing(), big.GetOrigin(), big.GetDirection())
    fmask = Image((rng.random(f.array.shape) > 0.3).astype(np.uint8), f.GetSpacing())
    mmask = Image((rng.random(mv.array.shape) > 0.2).astype(np.uint8), mv.GetSpacing(), mv.GetOrigin(), mv.GetDirection())
    init = linear.centered_transform_initializer(f, mv)
    m = linear.make_model("affine")
    p = m.identity() + 0.02 * rng.standard_normal(m.n)
    A, b = init.matrix @ m.matrix(p), init.matrix @ m.offset(p) + init.offset
    df, dm, dfm, dmm = (engine.to_device(i) for i in (f, mv, fmask, mmask))
    fb, mb = linear.mattes_bins(*engine.minmax(df)), linear.mattes_bins(*engine.minmax(dm))
    assert np.allclose(fb, linear.mattes_bins(f.array.min(), f.array.max())) and np.allclose(mb, linear.mattes_bins(mv.array.min(), mv.array.max()))
    for fm, mm, dfm_, dmm_, stride in ((None, None, None, None, 1), (fmask, mmask, dfm, dmm, 2)):
        hist, count = engine.linreg_mattes_histogram(df, dm, A, b, fb, mb, 50, dfm_, dmm_, stride)
        exp_hist, exp_count = ref.linreg_mattes(f, mv, A, b, init.matrix, m.center, fb, mb, 50, None, fm, mm, stride)
        assert count == exp_count and np.allclose(hist, exp_hist, rtol=0, atol=exp_count * 2.0 ** -32)
        again, _ = engine.linreg_mattes_histogram(df, dm, A, b, fb, mb, 50, dfm_, dmm_, stride)
        assert np.array_equal(hist, again)  # fixed-point integer atomics: deterministic
        _, table, _ = linear.mattes_value_and_table(exp_hist)
        sums = engine.linreg_mattes_derivative(df, dm, A, b, init.matrix, m.center, fb, mb, table, dfm_, dmm_, stride)
        _, _, exp_sums = ref.linreg_mattes(f, mv, A, b, init.matrix, m.center, fb, mb, 50, table, fm, mm, stride)
        assert np.allclose(sums, exp_sums, rtol=1e-9, atol=1e-9 * np.abs(exp_sums).max())
    fixed = blob((48, 40, 32), (24.0, 20.0, 16.0))
    shifted = blob((48, 40, 32), (26.0, 18.5, 17.0))
    moving = Image((1000.0 - shifted.array * 0.8).astype(np.float32), shifted.GetSpacing())
    for optimiser in ("gradient_descent_line_search", "lbfgsb"):
        _, tfm = linear.linear_registration(fixed, moving, reg_method="translation", metric="mattes_mi", optimiser=optimiser, shrink_factors=[2, 1],
                                            smooth_sigmas=[1, 0], sampling_rate=0.5, number_of_iterations=50, default_value=1000)
        pt = np.array((24.0, 20.0, 16.0))
        for part in tfm.flatten():
            pt = np.array(part.TransformPoint(pt))
        assert np.allclose(pt, (26.0, 18.5, 17.0), atol=0.5), (optimiser, pt)
    with pytest.raises(NotImplementedError):
        linear.linear_registration(fixed, moving, metric="joint_hist_mi")
    with pytest.raises(RuntimeError):
        linear.linear_registration(fixed, Image(np.zeros((8, 8, 8), np.float32)), metric="mattes_mi", shrink_factors=[1], smooth_sigmas=[0])


def test_get_com(engine):
    """label/utils.py:61-84: scipy.ndimage.center_of_mass in array order, or the physical point."""
    from platipy_b200 import label_utils as lu

    arr = np.zeros((20, 30, 40), np.uint8)
    arr[4:11, 10:25, 7:30] = 1
    arr[12:15, 3:9, 30:38] = 1
    ang = 0.3
    rot = (np.cos(ang), -np.sin(ang), 0, np.sin(ang), np.cos(ang), 0, 0, 0, 1.0)
    lab = Image(arr, (0.9, 1.1, 2.5), (5.0, -3.0, 10.0), rot)
    exp = ndi.center_of_mass(arr)
    assert np.allclose(lu.get_com(lab, as_int=False), exp, rtol=1e-12)
    assert lu.get_com(lab) == [int(v) for v in exp]
    d = np.asarray(rot).reshape(3, 3)
    phys = np.asarray(lab.GetOrigin()) + d @ (np.asarray(lab.GetSpacing()) * np.asarray(exp[::-1]))
    assert np.allclose(lu.get_com(lab, real_coords=True), phys, rtol=1e-12)
    assert np.allclose(lu.get_com(engine.to_device(lab), as_int=False), exp, rtol=1e-12)


def test_gpu_matches_rows3_golden(engine):
    """The committed fixture of these rows (tests/golden/rows3_small.npz, generated by tests/golden/make_golden.py from the oracle)."""
    import os

    from platipy_b200 import fusion, linear

    here = os.path.dirname(os.path.abspath(__file__))
    g, r3 = np.load(os.path.join(here, "golden", "demons_small.npz")), np.load(os.path.join(here, "golden", "rows3_small.npz"))
    sp, og = tuple(g["spacing"]), tuple(g["origin"])
    F, M, L = Image(g["fixed"], sp, og), Image(g["moving"], sp, og), Image(g["label"], sp, og)
    dL = engine.to_device(L)
    assert np.array_equal(engine.to_host(engine.signed_maurer_distance_map(dL)).array, r3["maurer"])
    assert np.array_equal(engine.to_host(engine.label_contour(dL)).array, r3["contour"])
    assert np.array_equal(engine.to_host(engine.binary_dilate(dL, ball_offsets((2, 1, 1)))).array, r3["dilated"])
    assert np.array_equal(engine.to_host(engine.binary_erode(dL, ball_offsets((2, 1, 1)))).array, r3["eroded"])
    assert np.array_equal(gen.convert_mask_to_reg_structure(L, 3).array, r3["reg_structure"])
    shifted, _, dvf = gen.generate_field_shift(L, (2.5, -1.9, 1.8), 2)
    assert np.array_equal(shifted.array, r3["shift_mask"]) and np.array_equal(dvf.array, r3["shift_dvf"])
    params = {"patch_window_mm": 12, "resampled_voxel_size_mm": 3, "correlation_function": lambda x: x + 1}
    assert np.allclose(fusion.compute_weight_map(F, M, "patch_correlation", params).array, r3["patch_weight"], rtol=1e-5, atol=1e-6)
    dF, dM = engine.to_device(F), engine.to_device(M)
    sums = engine.linreg_correlation(dF, dM, r3["lin_matrix"], r3["lin_offset"], np.eye(3), np.array(og), None, None, 3)
    assert sums[0] == r3["corr_sums"][0] and np.allclose(sums, r3["corr_sums"], rtol=1e-9, atol=1e-9 * np.abs(r3["corr_sums"]).max())
    fb, mb = tuple(r3["mattes_bins"][:2]), tuple(r3["mattes_bins"][2:])
    assert np.allclose(linear.mattes_bins(*engine.minmax(dF)), fb) and np.allclose(linear.mattes_bins(*engine.minmax(dM)), mb)
    hist, count = engine.linreg_mattes_histogram(dF, dM, r3["lin_matrix"], r3["lin_offset"], fb, mb, 50, None, None, 3)
    assert count == float(r3["mattes_count"]) and np.allclose(hist, r3["mattes_hist"], rtol=0, atol=count * 2.0 ** -32)
    other = Image(r3["dilated"], sp, og)
    sm = cmp.compute_surface_metrics(L, other)
    assert sorted(sm) == [str(k) for k in r3["surface_metric_names"]]
    assert np.allclose([float(sm[k]) for k in sorted(sm)], r3["surface_metric_values"], rtol=1e-10)
    assert [int(v) for v in cmp.compute_apl(L, other, 1.0)] == [int(v) for v in r3["apl"]]
