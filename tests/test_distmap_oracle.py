"""CPU tests of the rows after the hot path that need distance maps:

* the oracle's restatement of SignedMaurerDistanceMap / LabelContour / BinaryDilate / BinaryErode is pinned against scipy
  (exact Euclidean distance transform, binary erosion / dilation);
* the CUDA kernels of platipy_b200/csrc/distmap_kernels.cuh, run under the serial host emulation of tests/emu (this
  container has no GPU), agree with the oracle BIT FOR BIT -- indexing, the single-precision Voronoi arithmetic, the
  sign / sqrt epilogue.  The -m gpu tests (test_gpu_zz_generation.py) repeat the comparison with the real launches;
* the restated generators behave (a shift moves the centre of mass, an expansion grows the label ...);
* the host statistics of iterative atlas removal pick out a deliberately wrong atlas.
"""
import ctypes as C
import os

import numpy as np
import pytest
import scipy.ndimage as ndi

from oracle import generation_ref as gref
from oracle import itk_oracle as orc
from platipy_b200.sitk_compat import Image

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _blobs(shape, seed, level=0.02, sigma=2.5):
    r = np.random.default_rng(seed)
    return (ndi.gaussian_filter(r.standard_normal(shape), sigma) > level).astype(np.uint8)


def _P(a):
    return a.ctypes.data_as(C.c_void_p)


# ---- the oracle against scipy -----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape,spacing", [((20, 33, 41), (1.0, 1.0, 1.0)), ((17, 40, 23), (0.9, 1.3, 2.5)), ((1, 30, 30), (1.0, 1.0, 1.0))])
def test_oracle_maurer_is_the_euclidean_distance_to_the_contour(shape, spacing):
    m = _blobs(shape, 11)
    fg = m != 0
    border = fg & ~ndi.binary_erosion(fg, structure=np.ones((3, 3, 3)), border_value=1)
    edt = ndi.distance_transform_edt(~border, sampling=spacing[::-1])
    d = orc.signed_maurer_distance_map(m, spacing)  # SimpleITK defaults: negative inside, distance (not squared), image spacing
    assert d.dtype == np.float32
    assert np.all(d[border] == 0)
    assert np.allclose(d, np.where(fg, -edt, edt), rtol=1e-6, atol=1e-5)
    d2 = orc.signed_maurer_distance_map(m, spacing, inside_is_positive=True, squared_distance=True)
    assert np.allclose(d2, np.where(fg, edt ** 2, -(edt ** 2)), rtol=1e-6, atol=1e-4)
    dv = orc.signed_maurer_distance_map(m, spacing, use_image_spacing=False)
    assert np.allclose(np.abs(dv), ndi.distance_transform_edt(~border), rtol=1e-6, atol=1e-5)


def test_oracle_contour_and_morphology_against_scipy():
    m = _blobs((18, 30, 34), 5)
    fg = m != 0
    cross = ndi.generate_binary_structure(3, 1)
    assert np.array_equal(orc.label_contour(m, False), (fg & ~ndi.binary_erosion(fg, structure=cross, border_value=1)).astype(np.uint8))
    assert np.array_equal(orc.label_contour(m, True), (fg & ~ndi.binary_erosion(fg, structure=np.ones((3, 3, 3)), border_value=1)).astype(np.uint8))
    lab = (m * (1 + (np.arange(34)[None, None, :] > 17))).astype(np.uint8)  # two labels touching each other
    c = orc.label_contour(lab, False)
    assert set(np.unique(c)) <= {0, 1, 2} and np.all(c[:, :, 17][lab[:, :, 17] == 1] == 1)  # the seam is contour on both sides
    for radius in ((2, 2, 1), (1, 0, 3), (0, 0, 0)):
        offs = gref.ball(radius)
        st = np.zeros((2 * radius[2] + 1, 2 * radius[1] + 1, 2 * radius[0] + 1), bool)
        st[offs[:, 2] + radius[2], offs[:, 1] + radius[1], offs[:, 0] + radius[0]] = True
        assert np.array_equal(orc.binary_morph(m, offs, True), ndi.binary_dilation(fg, structure=st).astype(np.uint8))
        assert np.array_equal(orc.binary_morph(m, offs, False), ndi.binary_erosion(fg, structure=st, border_value=1).astype(np.uint8))
        assert np.array_equal(orc.binary_morph(m, offs, False, boundary_to_foreground=False),
                              ndi.binary_erosion(fg, structure=st, border_value=0).astype(np.uint8))
    # values other than the foreground value 1 are left alone
    odd = m.copy()
    odd[0, 0, 0] = 7
    assert orc.binary_morph(odd, gref.ball((1, 1, 1)), False)[0, 0, 0] == 7


# ---- the CUDA kernels (host emulation) against the oracle: bit-exact --------------------------------------------------------
@pytest.mark.parametrize("shape,spacing", [((20, 33, 41), (1.0, 1.0, 1.0)), ((17, 40, 23), (0.9, 1.3, 2.5)), ((5, 7, 64), (0.7, 0.7, 3.0)),
                                           ((1, 9, 11), (1.0, 2.0, 3.0))])
def test_emulated_maurer_kernels_bit_exact(emu, shape, spacing):
    m = _blobs(shape, 3)
    nz, ny, nx = shape
    sp = np.array(spacing, np.float64)
    for inside_pos in (0, 1):
        for squared in (0, 1):
            for use_sp in (0, 1):
                exp = orc.signed_maurer_distance_map(m, spacing, inside_pos, squared, use_sp)
                out = np.empty(shape, np.float32)
                emu.emu_signed_maurer(_P(m), nx, ny, nz, _P(sp), inside_pos, squared, use_sp, _P(out), C.c_uint(3), C.c_uint(64))
                assert np.array_equal(exp.view(np.uint32), out.view(np.uint32)), (inside_pos, squared, use_sp)


def test_emulated_maurer_edge_cases(emu):
    sp = np.ones(3)
    for m in (np.zeros((4, 5, 6), np.uint8), np.ones((4, 5, 6), np.uint8)):  # no contour anywhere: +-sqrt(FLT_MAX), as ITK leaves it
        exp = orc.signed_maurer_distance_map(m)
        out = np.empty(m.shape, np.float32)
        emu.emu_signed_maurer(_P(m), 6, 5, 4, _P(sp), 0, 0, 1, _P(out), C.c_uint(2), C.c_uint(32))
        assert np.array_equal(exp.view(np.uint32), out.view(np.uint32))
        assert np.all(np.abs(out) == np.sqrt(np.float32(np.finfo(np.float32).max)))
    one = np.zeros((6, 6, 6), np.uint8)
    one[2, 3, 4] = 9  # any non-zero value is object; a single voxel is its own contour
    exp = orc.signed_maurer_distance_map(one, (1, 1, 2))
    out = np.empty(one.shape, np.float32)
    sp2 = np.array([1.0, 1.0, 2.0])
    emu.emu_signed_maurer(_P(one), 6, 6, 6, _P(sp2), 0, 0, 1, _P(out), C.c_uint(7), C.c_uint(32))
    assert np.array_equal(exp.view(np.uint32), out.view(np.uint32)) and out[2, 3, 4] == 0 and out[0, 3, 4] == 4.0


def test_emulated_contour_morphology_and_elementwise_kernels(emu):
    shape = (14, 26, 37)
    nz, ny, nx = shape
    m = _blobs(shape, 8)
    lab = (m * (1 + (np.arange(nx)[None, None, :] > nx // 2))).astype(np.uint8)
    for fully in (0, 1):
        out = np.empty_like(lab)
        emu.emu_label_contour(_P(lab), nx, ny, nz, fully, _P(out), C.c_uint(5), C.c_uint(32))
        assert np.array_equal(out, orc.label_contour(lab, fully))
    odd = m.copy()
    odd[3, 3, 3] = 5
    for radius in ((2, 1, 1), (0, 3, 0)):
        offs = np.ascontiguousarray(gref.ball(radius))
        for dilate in (0, 1):
            for bfg in (0, 1):
                out = np.empty_like(odd)
                emu.emu_binary_morph(_P(odd), nx, ny, nz, _P(offs), len(offs), dilate, bfg, _P(out), C.c_uint(4), C.c_uint(64))
                assert np.array_equal(out, orc.binary_morph(odd, offs, bool(dilate), bool(bfg))), (radius, dilate, bfg)
    a, b = _blobs(shape, 1) * 255, _blobs(shape, 2) * 3
    a, b = a.astype(np.uint8), b.astype(np.uint8)
    for op, fn in enumerate((np.bitwise_or, np.bitwise_and, np.add, np.bitwise_xor)):
        out = np.empty_like(a)
        emu.emu_u8_binary_op(_P(a), _P(b), op, _P(out), C.c_size_t(a.size), C.c_uint(3), C.c_uint(32))
        assert np.array_equal(out, fn(a, b))
    n = m.size
    rng = np.random.default_rng(0)
    field = rng.standard_normal((3,) + shape)
    out = np.empty_like(field)
    emu.emu_mask_f64(_P(field), _P(m), C.c_size_t(n), 3, C.c_double(0.0), _P(out), C.c_uint(3), C.c_uint(32))
    assert np.array_equal(out, np.where(m[None] != 0, field, 0.0))
    emu.emu_divide_f64(_P(field), C.c_double(3.7), _P(out), C.c_size_t(field.size), C.c_uint(3), C.c_uint(32))
    assert np.array_equal(out, field / 3.7)
    emu.emu_constant_field(_P(m), C.c_size_t(n), C.c_double(1.5), C.c_double(-2.0), C.c_double(0.25), _P(out), C.c_uint(3), C.c_uint(32))
    assert np.array_equal(out, np.where(m[None] != 0, np.array([1.5, -2.0, 0.25])[:, None, None, None], 0.0))
    emu.emu_constant_field(None, C.c_size_t(n), C.c_double(1.5), C.c_double(-2.0), C.c_double(0.25), _P(out), C.c_uint(3), C.c_uint(32))
    assert np.all(out[0] == 1.5) and np.all(out[1] == -2.0) and np.all(out[2] == 0.25)


@pytest.mark.parametrize("where,clip", [(("z", "inf"), (2, 1)), (("z", "sup"), (2, 0)), (("y", "post"), (1, 0)), (("y", "ant"), (1, 1)),
                                        (("x", "left"), (0, 0)), (("x", "right"), (0, 1)), (False, (-1, 1))])
def test_emulated_radial_bend_field_matches_the_numpy_restatement(emu, where, clip):
    shape = (12, 18, 22)
    nz, ny, nx = shape
    body = Image(_blobs(shape, 4, level=-0.05), (1.0, 1.2, 2.0))
    image = Image(np.zeros(shape, np.float32), (1.0, 1.2, 2.0))
    ref_zyx = (5, 9, 12)
    axis_zyx = [0.3, -0.2, -1.0]
    _, _, exp = gref.generate_field_radial_bend(image, body, ref_zyx, axis_zyx, 0.13, where, gaussian_smooth=0)
    axis = np.array(axis_zyx)
    axis = (axis / np.linalg.norm(axis))[::-1]
    out = np.empty((3,) + shape)
    emu.emu_radial_bend(_P(body.array), nx, ny, nz, ref_zyx[2], ref_zyx[1], ref_zyx[0], C.c_double(axis[0]), C.c_double(axis[1]), C.c_double(axis[2]),
                        C.c_double(0.13), clip[0], clip[1], _P(out), C.c_uint(3), C.c_uint(64))
    assert np.array_equal(np.moveaxis(out, 0, -1), exp.array)
    assert np.abs(out).max() > 0.5


# ---- the restated generators behave -----------------------------------------------------------------------------------------
def _ellipsoid(shape=(24, 40, 36), spacing=(1.0, 1.2, 1.5)):
    zz, yy, xx = np.mgrid[: shape[0], : shape[1], : shape[2]]
    m = (((xx - 18) / 8.0) ** 2 + ((yy - 20) / 9.0) ** 2 + ((zz - 12) / 5.0) ** 2 < 1).astype(np.uint8)
    return Image(m, spacing)


def _com(a):
    return np.array(np.where(a)).mean(axis=1)


def test_restated_generators_behave():
    mask = _ellipsoid()
    n0 = int(mask.array.sum())
    shifted, tfm, dvf = gref.generate_field_shift(mask, (3, -2.4, 4), gaussian_smooth=1)
    d = _com(shifted.array) - _com(mask.array)  # (z, y, x) voxels; spacing (x 1.0, y 1.2, z 1.5): expected about (2, -2, 4)
    assert d[0] > 1.0 and d[1] < -1.0 and d[2] > 2.5
    assert dvf.is_vector and dvf.array.dtype == np.float64 and tfm.GetDisplacementField() is not None
    contracted, _, _ = gref.generate_field_asymmetric_contract(mask, (0, 6, 0), gaussian_smooth=1)
    extended, _, _ = gref.generate_field_asymmetric_extend(mask, (0, 6, 0), gaussian_smooth=1)
    assert contracted.array.sum() < n0 < extended.array.sum()
    grown, _, _ = gref.generate_field_expand(mask, expand=3, gaussian_smooth=1)
    shrunk, _, _ = gref.generate_field_expand(mask, expand=(-3, -2.4, 0), gaussian_smooth=1)
    assert shrunk.array.sum() < n0 < grown.array.sum()
    reg = gref.convert_mask_to_reg_structure(mask, expansion=3)
    assert reg.array.dtype == np.float64 and reg.array.max() == 1.0 and reg.array.min() == 0.0
    assert np.all(reg.array[mask.array == 1] > 0)  # dilated by 3 mm first: every original voxel is strictly inside
    dm = gref.convert_mask_to_distance_map(mask, normalise=True)
    assert dm.array.dtype == np.float32 and dm.array.max() == 1.0


def test_distance_to_reference_of_a_dilated_label():
    mask = _ellipsoid()
    bigger = gref.binary_dilate(mask, (2, 2, 1))
    v = gref.evaluate_distance_to_reference(mask, bigger)
    assert v.ndim == 1 and len(v) == int((orc.label_contour(mask.array) == 1).sum())
    assert v.min() >= 1.0 and v.max() <= 3.1  # the test surface lies 1 .. 2 voxels (<= 3 mm) outside the reference surface
    assert len(gref.evaluate_distance_to_reference(mask, bigger, resample_factor=5)) == (len(v) + 4) // 5


# ---- iterative atlas removal: host statistics ---------------------------------------------------------------------------------
def test_iar_statistics_flag_the_outlier():
    from platipy_b200 import iar

    rng = np.random.default_rng(7)
    n_pts = 600
    base = np.abs(rng.normal(1.0, 0.3, n_pts))
    g_vals = [np.abs(base + rng.normal(0, 0.15, n_pts)) for _ in range(12)]
    g_vals.append(np.abs(base + rng.normal(0, 0.15, n_pts)) + np.where(np.arange(n_pts) < 200, 1.0, 0.0))  # a third of the surface is 1 mm off
    q = []
    for i, g in enumerate(g_vals):
        q.append(iar.q_value(iar.z_scores(g, g_vals[:i] + g_vals[i + 1:], "MAD")))
    limit = iar._outlier_limit(q, "IQR", 1.5, 10)
    assert q[-1] > 5 * limit and q[-1] > 5 * max(q[:-1])  # (the IQR rule on the 10 best may also flag a borderline inlier, as in the reference)
    q_std = [iar.q_value(iar.z_scores(g, g_vals[:i] + g_vals[i + 1:], "STD")) for i, g in enumerate(g_vals)]
    assert int(np.argmax(q_std)) == len(g_vals) - 1
    with pytest.raises(ValueError):
        iar.z_scores(g_vals[0], g_vals[1:], "median")
    with pytest.raises(AttributeError):
        iar.run_iar({}, "S", project_on_sphere=True)


# ---- compute_weight_map(vote_type="patch_correlation"): kernel (host emulation) against scipy.stats.pearsonr ------------------------
def test_emulated_patch_correlation_matches_pearsonr(emu):
    from scipy.stats import pearsonr

    rng = np.random.default_rng(12)
    shape = (9, 11, 13)
    nz, ny, nx = shape
    t = ndi.gaussian_filter(rng.standard_normal(shape), 1.0).astype(np.float32)
    m = (0.6 * t + 0.4 * ndi.gaussian_filter(rng.standard_normal(shape), 1.0)).astype(np.float32)
    t[:3, :4, :5] = 2.5  # a constant corner: pearsonr -> NaN -> 0
    for window in ((4, 4, 4), (5, 3, 2), (1, 1, 2), (8, 8, 8)):  # (z, y, x) like window_box_im
        out = np.empty(shape, np.float64)
        emu.emu_patch_correlation(_P(t), _P(m), nx, ny, nz, window[2], window[1], window[0], _P(out), C.c_uint(3), C.c_uint(64))
        exp = np.empty(shape, np.float64)
        for z, y, x in np.ndindex(*shape):
            sl = tuple(slice(max(c - (w - 1) // 2, 0), min(c + w // 2, n - 1) + 1) for c, w, n in zip((z, y, x), window, shape))
            a, b = t[sl].ravel().astype(np.float64), m[sl].ravel().astype(np.float64)
            if (a == a[0]).all() or (b == b[0]).all():
                exp[z, y, x] = 0.0
            else:
                exp[z, y, x] = pearsonr(a, b)[0]
        assert np.allclose(out, exp, rtol=0, atol=1e-12), window
        assert np.all(np.abs(out) <= 1.0) and (max(window) > 4 or out[0, 0, 0] == 0.0)
    vals = rng.standard_normal(1000)
    out = np.empty_like(vals)
    emu.emu_scale_shift_f64(_P(vals), C.c_size_t(vals.size), 1, C.c_double(2.0), C.c_double(1.0), _P(out), C.c_uint(3), C.c_uint(32))
    assert np.array_equal(out, np.abs(vals) * 2.0 + 1.0)
    v32 = vals.astype(np.float32)
    o32 = np.empty_like(v32)
    emu.emu_scale_shift_f32(_P(v32), C.c_size_t(v32.size), 0, C.c_float(1.0), C.c_float(1e-5), _P(o32), C.c_uint(3), C.c_uint(32))
    assert np.array_equal(o32, v32 + np.float32(1e-5))


def test_patch_correlation_weight_map_chain_on_the_cpu(emu):
    """The whole vote as the product composes it -- resample both images, the correlation kernel, resample back, the correlation
    function, cast -- with the oracle standing in for the two resampling steps, against the reference-shaped restatement
    (padding, window views, pearsonr per patch)."""
    from oracle import platipy_ref as ref
    from platipy_b200.synth import synth_pair

    f, m = synth_pair((40, 36, 24), seed=5, spacing=(1.0, 1.0, 2.0))
    for fn in (lambda x: x + 1, abs):
        vp = dict(patch_window_mm=12, resampled_voxel_size_mm=3, correlation_function=fn)
        exp = ref.compute_weight_map(f, m, "patch_correlation", vp)
        t_res, m_res = ref.smooth_and_resample(f, isotropic_voxel_size_mm=3), ref.smooth_and_resample(m, isotropic_voxel_size_mm=3)
        window = [int(12 / i) for i in t_res.GetSpacing()[::-1]]
        nz, ny, nx = t_res.array.shape
        corr = np.empty(t_res.array.shape, np.float64)
        emu.emu_patch_correlation(_P(t_res.array), _P(m_res.array), nx, ny, nz, window[2], window[1], window[0], _P(corr), C.c_uint(4), C.c_uint(64))
        back = ref.resample(Image(corr, t_res.GetSpacing(), t_res.GetOrigin(), t_res.GetDirection()), f)
        got = fn(back).array.astype(np.float32)
        assert exp.array.dtype == np.float32 and np.allclose(got, exp.array, rtol=1e-6, atol=1e-7)
        assert exp.array.min() >= 0.0
