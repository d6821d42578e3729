"""CPU tests of linear_registration's host logic (platipy_b200/linear.py: transform parameterisations, scales from
physical shift, learning-rate estimation, convergence window, pyramid geometry) -- the optimiser is driven by the
oracle's numpy metric here (test infrastructure), on the GPU box by b200reg_linreg_meansq."""
import numpy as np
import pytest

from oracle import platipy_ref as ref
from platipy_b200 import linear
from platipy_b200 import sitk_compat as sk
from platipy_b200.sitk_compat import Image


def _fd(model, p, eps=1e-6):
    out = []
    for k in range(model.n):
        d = np.zeros(model.n)
        d[k] = eps
        out.append((model.matrix(p + d) - model.matrix(p - d)) / (2 * eps))
    return out


def test_matrix_derivatives_match_finite_differences():
    rng = np.random.default_rng(0)
    for name in ("rigid", "similarity", "affine", "scale", "ScaleVersor", "ScaleSkewVersor"):
        m = linear.make_model(name)
        p = m.identity() + 0.1 * rng.standard_normal(m.n)
        fd = _fd(m, p)
        got = dict(m.matrix_bases(p))
        for k in range(m.n):
            exp = fd[k]
            if k in got:
                assert np.allclose(got[k], exp, atol=1e-7), (name, k)
            else:
                assert np.allclose(exp, 0, atol=1e-9), (name, k)  # translation parameters do not move the matrix


def test_versor_update_composes_rotations():
    m = linear.make_model("rigid")
    p = m.identity()
    p = m.updated(p, np.array([0, 0, 0.3, 1.0, 2.0, 3.0]))  # angle 0.3 about z, translation added
    c, s = np.cos(0.3), np.sin(0.3)
    assert np.allclose(m.matrix(p), [[c, -s, 0], [s, c, 0], [0, 0, 1]])
    assert np.allclose(m.translation(p), [1, 2, 3])
    p2 = m.updated(p, np.array([0, 0, 0.2, 0, 0, 0]))
    c, s = np.cos(0.5), np.sin(0.5)
    assert np.allclose(m.matrix(p2), [[c, -s, 0], [s, c, 0], [0, 0, 1]])
    r = m.matrix(m.updated(p2, np.array([0.4, -0.2, 0.1, 0, 0, 0])))
    assert np.allclose(r @ r.T, np.eye(3), atol=1e-12) and np.isclose(np.linalg.det(r), 1.0)


def test_shrink_grid_and_initialiser():
    im = Image(np.zeros((32, 40, 64), np.float32), (1.0, 1.5, 2.5), (10.0, -5.0, 3.0))
    g = linear.shrink_grid(im, 8)
    assert g.GetSize() == (8, 5, 4) and g.GetSpacing() == (8.0, 12.0, 20.0)
    assert np.allclose(linear.image_center(g), linear.image_center(im))  # ShrinkImageFilter keeps the physical centre
    mv = Image(np.zeros((10, 10, 10), np.float32), (2.0, 2.0, 2.0), (100.0, 0.0, 0.0))
    t = linear.centered_transform_initializer(im, mv)
    assert np.allclose(t.TransformPoint(linear.image_center(im)), linear.image_center(mv))
    with pytest.raises(ValueError):
        linear.make_model("nonsense")
    sv, ssv = linear.make_model("ScaleVersor"), linear.make_model("ScaleSkewVersor")
    assert (sv.n, ssv.n) == (9, 15) and np.array_equal(sv.matrix(), np.eye(3)) and np.array_equal(ssv.matrix(), np.eye(3))
    p = ssv.identity()
    p[6:9], p[9:] = (1.1, 0.9, 1.2), (0.01, 0.02, 0.03, 0.04, 0.05, 0.06)
    assert np.allclose(ssv.matrix(p), [[1.1, 0.01, 0.02], [0.03, 0.9, 0.04], [0.05, 0.06, 1.2]])  # identity rotation + scale - 1 + skew


def test_scales_and_convergence_value():
    m = linear.make_model("similarity")
    corners = linear.image_corners(Image(np.zeros((50, 100, 200), np.float32)))
    sc = linear.estimate_scales(m, m.identity(), corners)
    assert np.allclose(sc[3:6], 1.0)  # a unit translation shifts every point by one unit
    far = np.linalg.norm(corners, axis=1).max()
    assert np.isclose(sc[6], far ** 2, rtol=1e-6)  # scaling about the origin moves the farthest corner by its distance
    assert sc[0] > 1e3
    assert linear.convergence_value([5.0] * 10) == pytest.approx(0.0, abs=1e-12)
    assert linear.convergence_value(list(np.linspace(10, 1, 10))) > 1e-2


def _blob_image(size, center, spacing=(1.0, 1.0, 1.0), sig=(7.0, 5.0, 4.0)):
    nx, ny, nz = size
    z, y, x = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    p = [x * spacing[0], y * spacing[1], z * spacing[2]]
    v = sum(((pi - ci) / si) ** 2 for pi, ci, si in zip(p, center, sig))
    return Image((1000.0 * np.exp(-0.5 * v)).astype(np.float32), spacing)


def test_metric_gradient_is_the_derivative_of_the_value():
    f = _blob_image((24, 20, 16), (12.0, 10.0, 8.0))
    # the moving grid is larger than the fixed one, so every sample stays inside the buffer and the sample count
    # does not depend on the parameters (ITK's derivative ignores that dependence as well)
    big = _blob_image((36, 32, 28), (19.5, 15.0, 14.5))
    mv = Image(big.array, big.GetSpacing(), (-6.0, -6.0, -6.0))
    init = linear.centered_transform_initializer(f, mv)
    for name in ("translation", "rigid", "similarity", "affine", "scaleversor", "scaleskewversor"):
        m = linear.make_model(name)
        rng = np.random.default_rng(3)
        p = m.identity() + 0.01 * rng.standard_normal(m.n)

        def value(q):
            acc = ref.linreg_meansq(f, mv, init.matrix @ m.matrix(q), init.matrix @ m.offset(q) + init.offset, init.matrix, m.center)
            return acc[0] / acc[1], acc

        v0, acc = value(p)
        g = m.gradient(acc, p)
        for k in range(m.n):
            d = np.zeros(m.n)
            d[k] = 1e-5
            fd = (value(p + d)[0] - value(p - d)[0]) / 2e-5
            assert np.isclose(g[k], fd, rtol=2e-3, atol=1e-3 * abs(g).max()), (name, k, g[k], fd)


def test_optimiser_recovers_a_translation_with_the_oracle_metric():
    f = _blob_image((24, 20, 16), (12.0, 10.0, 8.0))
    mv = _blob_image((24, 20, 16), (14.0, 8.5, 9.0))   # moving = fixed shifted by (+2, -1.5, +1)
    init = linear.centered_transform_initializer(f, mv)
    m = linear.make_model("translation")

    def evaluate(p):
        return ref.linreg_meansq(f, mv, init.matrix @ m.matrix(p), init.matrix @ m.offset(p) + init.offset, init.matrix, m.center, stride=2)

    hist = linear.optimise_level(m, evaluate, linear.image_corners(f), 1.0, 60)
    assert hist[-1] < 0.02 * hist[0]
    assert np.allclose(m.p, [2.0, -1.5, 1.0], atol=0.15)


def test_golden_section_and_line_search_optimiser():
    assert linear.golden_section(lambda x: (x - 1.7) ** 2, 0.0, 1.0, 5.0) == pytest.approx(1.7, abs=0.03)
    assert linear.golden_section(lambda x: (x - 0.2) ** 2, 0.0, 1.0, 5.0) == pytest.approx(0.2, abs=0.03)
    f = _blob_image((24, 20, 16), (12.0, 10.0, 8.0))
    mv = _blob_image((24, 20, 16), (14.0, 8.5, 9.0))
    init = linear.centered_transform_initializer(f, mv)
    m = linear.make_model("translation")

    def evaluate(p):
        return ref.linreg_meansq(f, mv, init.matrix @ m.matrix(p), init.matrix @ m.offset(p) + init.offset, init.matrix, m.center, stride=2)

    hist = linear.optimise_level_line_search(m, evaluate, linear.image_corners(f), 1.0, 15)
    assert hist[-1] < 0.01 * hist[0]
    assert np.allclose(m.p, [2.0, -1.5, 1.0], atol=0.1)


# ---- metric "correlation" (linear.py:141-146) --------------------------------------------------------------------------------
def _corr_value(f, mv, init, m, q, **kw):
    sums = ref.linreg_correlation(f, mv, init.matrix @ m.matrix(q), init.matrix @ m.offset(q) + init.offset, init.matrix, m.center, **kw)
    acc = linear.correlation_in_meansq_form(sums)
    return acc[0] / acc[1], acc


def test_correlation_value_and_gradient():
    f = _blob_image((24, 20, 16), (12.0, 10.0, 8.0))
    big = _blob_image((36, 32, 28), (19.5, 15.0, 14.5))
    mv = Image(big.array * 0.5 + 40.0, big.GetSpacing(), (-6.0, -6.0, -6.0))  # another intensity scale: correlation does not care
    init = linear.centered_transform_initializer(f, mv)
    m = linear.make_model("translation")
    v_aligned, _ = _corr_value(f, mv, init, m, np.array([0.0, -0.5, 0.0]))
    assert -1.0 <= v_aligned < -0.9
    # the value is Pearson's r squared (negated) of the sampled pairs
    sums = ref.linreg_correlation(f, mv, init.matrix, init.offset, init.matrix, m.center)
    n, sf, sm, sff, smm, sfm = sums[:6]
    r = (sfm - sf * sm / n) / np.sqrt((sff - sf * sf / n) * (smm - sm * sm / n))
    acc = linear.correlation_in_meansq_form(sums)
    assert acc[1] == n and np.isclose(acc[0] / n, -r * r, rtol=1e-12)
    for name in ("translation", "rigid", "similarity", "affine"):
        m = linear.make_model(name)
        rng = np.random.default_rng(3)
        p = m.identity() + 0.01 * rng.standard_normal(m.n)
        v0, acc = _corr_value(f, mv, init, m, p)
        g = m.gradient(acc, p)
        for k in range(m.n):
            d = np.zeros(m.n)
            d[k] = 1e-5
            fd = (_corr_value(f, mv, init, m, p + d)[0] - _corr_value(f, mv, init, m, p - d)[0]) / 2e-5
            assert np.isclose(g[k], fd, rtol=2e-3, atol=1e-3 * abs(g).max()), (name, k, g[k], fd)
    # degenerate sample sets: no valid sample, constant image
    assert not linear.correlation_in_meansq_form(np.zeros(42)).any()
    flat = Image(np.full_like(f.array, 3.0), f.GetSpacing())
    acc = linear.correlation_in_meansq_form(ref.linreg_correlation(flat, mv, init.matrix, init.offset, init.matrix, m.center))
    assert acc[0] == 0.0 and acc[1] > 0 and not acc[2:].any()


def test_optimiser_recovers_a_translation_with_the_correlation_metric():
    f = _blob_image((24, 20, 16), (12.0, 10.0, 8.0))
    shifted = _blob_image((24, 20, 16), (14.0, 8.5, 9.0))
    mv = Image(shifted.array * 2.0 + 100.0, shifted.GetSpacing())  # mean squares cannot align these; correlation can
    init = linear.centered_transform_initializer(f, mv)
    for run in (linear.optimise_level, linear.optimise_level_line_search):
        m = linear.make_model("translation")

        def evaluate(p):
            return linear.correlation_in_meansq_form(
                ref.linreg_correlation(f, mv, init.matrix @ m.matrix(p), init.matrix @ m.offset(p) + init.offset, init.matrix, m.center, stride=2))

        hist = run(m, evaluate, linear.image_corners(f), 1.0, 60)
        assert hist[-1] < -0.98 and hist[-1] < hist[0]
        assert np.allclose(m.p, [2.0, -1.5, 1.0], atol=0.2), (run.__name__, m.p)


def test_emulated_correlation_kernel_matches_the_numpy_sums(emu):
    """The CUDA kernel's per-sample code, run on the host (tests/emu), against the numpy restatement -- masks, stride, a rotated
    moving grid, samples falling outside the moving buffer."""
    import ctypes as C

    rng = np.random.default_rng(9)
    f = _blob_image((24, 20, 16), (12.0, 10.0, 8.0), spacing=(1.0, 1.2, 1.5))
    ang = 0.15
    rot = np.array([[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1.0]])
    big = _blob_image((30, 26, 20), (15.0, 13.5, 12.0), spacing=(1.1, 1.0, 1.4))
    mv = Image((big.array + rng.normal(0, 5, big.array.shape)).astype(np.float32), big.GetSpacing(), (-3.0, 2.0, -1.0), tuple(rot.reshape(9)))
    fmask = Image((rng.random(f.array.shape) > 0.3).astype(np.uint8), f.GetSpacing())
    mmask = Image((rng.random(mv.array.shape) > 0.2).astype(np.uint8), mv.GetSpacing(), mv.GetOrigin(), mv.GetDirection())
    init = linear.centered_transform_initializer(f, mv)
    m = linear.make_model("affine")
    p = m.identity() + 0.02 * rng.standard_normal(m.n)
    A, b = init.matrix @ m.matrix(p), init.matrix @ m.offset(p) + init.offset

    def geo(img):
        d = np.asarray(img.GetDirection(), np.float64).reshape(3, 3)
        i2p = d * np.asarray(img.GetSpacing())[None, :]
        return (np.array(img.GetSize(), np.int32), np.concatenate([np.asarray(img.GetOrigin(), np.float64), i2p.reshape(9), np.linalg.inv(i2p).reshape(9)]))

    (fs, fg), (ms, mg) = geo(f), geo(mv)
    pose = np.concatenate([A.reshape(9), b, init.matrix.T.reshape(9), m.center]).astype(np.float64)
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    for fm, mm, stride in ((None, None, 1), (fmask, None, 3), (fmask, mmask, 2)):
        grid, block = 3, 64
        partials = np.zeros((grid * block, 42))
        emu.emu_linreg_corr(P(f.array), P(mv.array), P(fm.array) if fm else None, P(mm.array) if mm else None, P(fs), P(fg), P(ms), P(mg), P(pose),
                            stride, P(partials), C.c_uint(grid), C.c_uint(block))
        got = partials.sum(axis=0)
        exp = ref.linreg_correlation(f, mv, A, b, init.matrix, m.center, fm, mm, stride)
        assert got[0] == exp[0] and 0 < got[0] < f.array.size / stride + 1
        assert np.allclose(got, exp, rtol=1e-9, atol=1e-6 * np.abs(exp).max()), (stride, np.abs(got - exp).max())


def test_lbfgsb_optimiser_with_the_oracle_metrics():
    """optimiser="lbfgsb" (linear.py:208-216): scipy's L-BFGS-B on the scaled parameters, for both metrics and for a
    transform with a versor part (parameters set directly, not composed)."""
    f = _blob_image((24, 20, 16), (12.0, 10.0, 8.0))
    shifted = _blob_image((24, 20, 16), (14.0, 8.5, 9.0))
    init = linear.centered_transform_initializer(f, shifted)
    for name, metric, mv in (("translation", "mean_squares", shifted), ("translation", "correlation", Image(shifted.array * 2.0 + 100.0, shifted.GetSpacing())),
                             ("rigid", "mean_squares", shifted)):
        m = linear.make_model(name)
        m.center = linear.image_center(f)

        def evaluate(p):
            args = (f, mv, init.matrix @ m.matrix(p), init.matrix @ m.offset(p) + init.offset, init.matrix, m.center)
            if metric == "correlation":
                return linear.correlation_in_meansq_form(ref.linreg_correlation(*args, stride=2))
            return ref.linreg_meansq(*args, stride=2)

        hist = linear.optimise_level_lbfgsb(m, evaluate, linear.image_corners(f), 1.0, 50)
        assert len(hist) <= 1024 and (min(hist) < -0.98 if metric == "correlation" else min(hist) < 0.01 * hist[0])
        assert np.allclose(m.translation(m.p), [2.0, -1.5, 1.0], atol=0.1), (name, metric, m.p)
        if name == "rigid":
            assert np.allclose(m.matrix(m.p), np.eye(3), atol=0.02)


def test_emulated_image_moments_and_center_of_gravity(emu):
    """alignment_registration(moments=True): the moments kernel (host emulation) against the numpy restatement; the centre of
    gravity of a Gaussian blob is its centre."""
    import ctypes as C

    ang = 0.2
    rot = (np.cos(ang), -np.sin(ang), 0, np.sin(ang), np.cos(ang), 0, 0, 0, 1.0)
    blob = _blob_image((30, 26, 20), (15.0, 11.0, 12.0), spacing=(1.1, 1.0, 1.4))
    img = Image(blob.array, blob.GetSpacing(), (-3.0, 2.0, -1.0), rot)
    d = np.asarray(rot).reshape(3, 3)
    geo = np.concatenate([np.asarray(img.GetOrigin()), (d * np.asarray(img.GetSpacing())[None, :]).reshape(9)])
    size = np.array(img.GetSize(), np.int32)
    grid, block = 3, 64
    partials = np.zeros((grid * block, 4))
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    emu.emu_image_moments(P(img.array), P(size), P(geo), P(partials), C.c_uint(grid), C.c_uint(block))
    got, exp = partials.sum(axis=0), ref.image_moments(img)
    assert np.allclose(got, exp, rtol=1e-12)
    cog = linear.center_of_gravity(got)
    assert np.allclose(cog, np.asarray(img.GetOrigin()) + d @ np.array([15.0, 11.0, 12.0]), atol=0.25)  # the blob is cut asymmetrically by the image border
    with pytest.raises(RuntimeError):
        linear.center_of_gravity(np.zeros(4))


# ---- metric "mattes_mi" (linear.py:145-146) -----------------------------------------------------------------------------------
def _mattes_acc(f, mv, init, m, q, fb, mb, **kw):
    A, b = init.matrix @ m.matrix(q), init.matrix @ m.offset(q) + init.offset
    hist, count = ref.linreg_mattes(f, mv, A, b, init.matrix, m.center, fb, mb, **kw)
    value, table, total = linear.mattes_value_and_table(hist)
    _, _, sums = ref.linreg_mattes(f, mv, A, b, init.matrix, m.center, fb, mb, table=table, **kw)
    return linear.mattes_in_meansq_form(value, count, total, sums)


def test_mattes_value_gradient_and_optimisers():
    f = _blob_image((24, 20, 16), (12.0, 10.0, 8.0))
    big = _blob_image((36, 32, 28), (19.5, 15.0, 14.5))
    mv = Image(1000.0 - big.array * 0.8, big.GetSpacing(), (-6.0, -6.0, -6.0))  # inverted contrast: mean squares and correlation^1 fail here
    init = linear.centered_transform_initializer(f, mv)
    fb, mb = linear.mattes_bins(f.array.min(), f.array.max()), linear.mattes_bins(mv.array.min(), mv.array.max())
    assert np.isclose(fb[0] * 46, float(f.array.max()) - float(f.array.min())) and np.isclose(fb[1], f.array.min() / fb[0] - 2)
    # histogram: partition of unity (every sample adds weight 1), marginal of the fixed axis = plain histogram of the fixed bins
    m = linear.make_model("translation")
    hist, count = ref.linreg_mattes(f, mv, init.matrix, init.offset, init.matrix, m.center, fb, mb)
    assert count == f.array.size and np.isclose(hist.sum(), count, rtol=1e-9) and hist.shape == (50, 50)
    assert not hist[:2].any() and not hist[-2:].any()  # the padding bins of the fixed axis stay empty
    value, table, total = linear.mattes_value_and_table(hist)
    assert value < -0.5 and np.isfinite(table).all()
    for name in ("translation", "rigid", "affine"):
        m = linear.make_model(name)
        rng = np.random.default_rng(3)
        p = m.identity() + 0.01 * rng.standard_normal(m.n)
        acc = _mattes_acc(f, mv, init, m, p, fb, mb)
        g = m.gradient(acc, p)
        for k in range(m.n):
            d = np.zeros(m.n)
            d[k] = 1e-5
            a1, a0 = _mattes_acc(f, mv, init, m, p + d, fb, mb), _mattes_acc(f, mv, init, m, p - d, fb, mb)
            fd = (a1[0] / a1[1] - a0[0] / a0[1]) / 2e-5
            assert np.isclose(g[k], fd, rtol=2e-3, atol=1e-3 * abs(g).max()), (name, k, g[k], fd)
    # registration of the inverted-contrast pair
    shifted = _blob_image((24, 20, 16), (14.0, 8.5, 9.0))
    inv = Image(1000.0 - shifted.array * 0.8, shifted.GetSpacing())
    init = linear.centered_transform_initializer(f, inv)
    mb = linear.mattes_bins(inv.array.min(), inv.array.max())
    for run in (linear.optimise_level, linear.optimise_level_line_search, linear.optimise_level_lbfgsb):
        m = linear.make_model("translation")
        hist = run(m, lambda p: _mattes_acc(f, inv, init, m, p, fb, mb, stride=2), linear.image_corners(f), 1.0, 60)
        assert min(hist) < hist[0] - 0.2
        assert np.allclose(m.p, [2.0, -1.5, 1.0], atol=0.3), (run.__name__, m.p)
    with pytest.raises(RuntimeError):
        linear.mattes_bins(3.0, 3.0)
    assert not linear.mattes_in_meansq_form(0.0, 0, 0.0, np.zeros(12)).any()


def test_emulated_mattes_kernels_match_the_numpy_restatement(emu):
    import ctypes as C

    rng = np.random.default_rng(9)
    f = _blob_image((24, 20, 16), (12.0, 10.0, 8.0), spacing=(1.0, 1.2, 1.5))
    ang = 0.15
    rot = np.array([[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1.0]])
    big = _blob_image((30, 26, 20), (15.0, 13.5, 12.0), spacing=(1.1, 1.0, 1.4))
    mv = Image((900.0 - big.array + rng.normal(0, 5, big.array.shape)).astype(np.float32), big.GetSpacing(), (-3.0, 2.0, -1.0), tuple(rot.reshape(9)))
    fmask = Image((rng.random(f.array.shape) > 0.3).astype(np.uint8), f.GetSpacing())
    mmask = Image((rng.random(mv.array.shape) > 0.2).astype(np.uint8), mv.GetSpacing(), mv.GetOrigin(), mv.GetDirection())
    init = linear.centered_transform_initializer(f, mv)
    m = linear.make_model("affine")
    p = m.identity() + 0.02 * rng.standard_normal(m.n)
    A, b = init.matrix @ m.matrix(p), init.matrix @ m.offset(p) + init.offset
    fb, mb = linear.mattes_bins(f.array.min(), f.array.max()), linear.mattes_bins(mv.array.min(), mv.array.max())

    def geo(img):
        d = np.asarray(img.GetDirection(), np.float64).reshape(3, 3)
        i2p = d * np.asarray(img.GetSpacing())[None, :]
        return (np.array(img.GetSize(), np.int32), np.concatenate([np.asarray(img.GetOrigin(), np.float64), i2p.reshape(9), np.linalg.inv(i2p).reshape(9)]))

    (fs, fg), (ms, mg) = geo(f), geo(mv)
    pose = np.concatenate([A.reshape(9), b, init.matrix.T.reshape(9), m.center]).astype(np.float64)
    bins = np.array([fb[0], fb[1], mb[0], mb[1]])
    P = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None
    for fm, mm, stride in ((None, None, 1), (fmask, mmask, 2)):
        grid, block, nb, replicas = 5, 64, 50, 3  # the blocks spread their atomics over `replicas` histogram copies
        hist_fp = np.zeros(replicas * nb * nb + 1, np.uint64)
        emu.emu_linreg_mattes(P(f.array), P(mv.array), P(fm.array) if fm else None, P(mm.array) if mm else None, P(fs), P(fg), P(ms), P(mg), P(pose),
                              stride, nb, P(bins), replicas, P(hist_fp), None, None, C.c_uint(grid), C.c_uint(block))
        exp_hist, exp_count = ref.linreg_mattes(f, mv, A, b, init.matrix, m.center, fb, mb, nb, None, fm, mm, stride)
        per_replica = hist_fp[:-1].reshape(replicas, nb, nb)
        assert all(per_replica[r].any() for r in range(replicas))
        got_hist = per_replica.sum(axis=0).astype(np.float64) / 2.0 ** 32
        assert float(hist_fp[-1]) == exp_count and np.allclose(got_hist, exp_hist, rtol=0, atol=exp_count * 2.0 ** -32)
        _, table, _ = linear.mattes_value_and_table(exp_hist)
        partials = np.zeros((grid * block, 12))
        emu.emu_linreg_mattes(P(f.array), P(mv.array), P(fm.array) if fm else None, P(mm.array) if mm else None, P(fs), P(fg), P(ms), P(mg), P(pose),
                              stride, nb, P(bins), replicas, None, P(np.ascontiguousarray(table)), P(partials), C.c_uint(grid), C.c_uint(block))
        _, _, exp_sums = ref.linreg_mattes(f, mv, A, b, init.matrix, m.center, fb, mb, nb, table, fm, mm, stride)
        assert np.allclose(partials.sum(axis=0), exp_sums, rtol=1e-9, atol=1e-9 * np.abs(exp_sums).max())
