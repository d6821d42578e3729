"""CPU checks of the oracle itself.  The reference holds no golden vectors for this path and SimpleITK
cannot be installed here (PARITY UNPINNED, see oracle/itk_oracle.c), so the oracle is pinned as far as
possible against independent implementations (scipy) and analytic properties, and against the one
acceptance criterion the reference's tests own: the Dice > 0.99 sphere-phantom check of
platipy/imaging/tests/test_cardiac.py:35-71,142 (restated for the Demons stage)."""
import numpy as np
import pytest
import scipy.ndimage as ndi
import scipy.special as sps

from oracle import itk_oracle as orc
from oracle import platipy_ref as ref
from platipy_b200 import sitk_compat as sk
from platipy_b200.sitk_compat import Image
from platipy_b200.synth import insert_sphere, smooth_random_dvf, synth_pair

IDENT = (1, 0, 0, 0, 1, 0, 0, 0, 1)


def test_gaussian_operator_matches_exact_bessel_and_itk_radii():
    # radii recorded in SURVEY 2.2 N1 (maxError 0.01, 1 mm spacing): sigma 1 -> 3, 2 -> 5, 4 -> 10, 8 -> 21
    for sigma, radius in [(1, 3), (2, 5), (4, 10), (8, 21)]:
        k = orc.gaussian_operator(sigma * sigma, 0.01, 1000)
        assert (len(k) - 1) // 2 == radius
        assert abs(k.sum() - 1.0) < 1e-15
        n = np.arange(-radius, radius + 1)
        exact = sps.ive(np.abs(n), sigma * sigma)
        exact /= exact.sum()
        # ITK uses the Numerical-Recipes polynomial / Miller-recurrence Bessel functions: ~1e-8 for small t,
        # a few 1e-6 at t = 64 (the recurrence starts only moderately above the argument)
        assert np.abs(k - exact).max() < (5e-8 if sigma < 8 else 1e-5)
    # Demons kernels: update field (var 1, maxError 0.1) and 1.5 mm regularisation at 1 mm spacing -> radius 2
    assert len(orc.gaussian_operator(1.0, 0.1, 30)) == 5
    assert len(orc.gaussian_operator(2.25, 0.1, 30)) == 5
    # maximum kernel width truncates: width 2 -> stops once the one-sided length exceeds it
    assert len(orc.gaussian_operator(64.0, 0.01, 2)) == 2 * 2 + 1


def test_discrete_gaussian_matches_scipy_separable_clamp():
    rng = np.random.default_rng(0)
    a = rng.normal(size=(12, 17, 21)).astype(np.float32)
    g = orc.make_geom((21, 17, 12), (1.0, 0.8, 2.0), (0, 0, 0), IDENT)
    var = (2.0, 1.5, 5.0)  # mm^2 along x, y, z
    out = orc.discrete_gaussian_f32(a, g, var, 32, 0.01)
    exp = a.astype(np.float64)
    for axis_np, ax in [(0, 2), (1, 1), (2, 0)]:  # z, y, x order with float32 intermediates
        k = orc.gaussian_operator(var[ax] / g.spacing[ax] ** 2, 0.01, 32)
        exp = ndi.correlate1d(exp, k, axis=axis_np, mode="nearest").astype(np.float32).astype(np.float64)
    assert np.abs(out - exp).max() < 1e-6


def test_linear_and_nn_resample_match_scipy_in_the_interior():
    rng = np.random.default_rng(1)
    a = rng.normal(size=(10, 12, 14)).astype(np.float32)
    gin = orc.make_geom((14, 12, 10), (1.0, 1.0, 1.0), (0, 0, 0), IDENT)
    # affine: small rotation about z + translation
    th = 0.1
    M = np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1.0]])
    off = np.array([0.7, -0.4, 0.3])
    out = orc.resample_scalar(a, gin, gin, [("affine", M, off)], 2, -5.0)
    zz, yy, xx = np.meshgrid(np.arange(10), np.arange(12), np.arange(14), indexing="ij")
    P = np.stack([xx, yy, zz], -1).astype(np.float64) @ M.T + off
    coords = [P[..., 2], P[..., 1], P[..., 0]]
    exp = ndi.map_coordinates(a.astype(np.float64), coords, order=1, mode="constant", cval=np.nan)
    inside = np.all([(P[..., 0] >= 0), (P[..., 0] <= 13), (P[..., 1] >= 0), (P[..., 1] <= 11), (P[..., 2] >= 0), (P[..., 2] <= 9)], axis=0)
    assert inside.sum() > 500
    assert np.abs(out[inside] - exp[inside]).max() < 1e-5
    # outside the buffer (beyond the half-voxel margin) -> default value
    far = (P[..., 0] < -0.5) | (P[..., 0] >= 13.5) | (P[..., 1] < -0.5) | (P[..., 1] >= 11.5)
    assert np.all(out[far] == -5.0)
    # nearest neighbour on labels: round-half-up
    lab = rng.integers(0, 5, size=a.shape).astype(np.uint8)
    outl = orc.resample_scalar(lab, gin, gin, [("affine", M, off)], 1, 0)
    idx = np.floor(P + 0.5).astype(int)
    ok = np.all([(P[..., 0] >= -0.5), (P[..., 0] < 13.5), (P[..., 1] >= -0.5), (P[..., 1] < 11.5), (P[..., 2] >= -0.5), (P[..., 2] < 9.5)], axis=0)
    expl = np.zeros_like(lab)
    expl[ok] = lab[idx[..., 2][ok], idx[..., 1][ok], idx[..., 0][ok]]
    assert np.array_equal(outl, expl)


def test_identity_resample_and_integer_cast():
    rng = np.random.default_rng(2)
    a = (rng.normal(size=(6, 7, 8)) * 100).astype(np.int16)
    g = orc.make_geom((8, 7, 6), (1.0, 0.5, 2.0), (320, -52, 60), IDENT)
    assert np.array_equal(orc.resample_scalar(a, g, g, [], 1, 0), a)
    # linear interpolation onto a half-voxel shifted grid truncates toward zero for integer outputs
    g2 = orc.make_geom((8, 7, 6), (1.0, 0.5, 2.0), (320 + 0.5, -52, 60), IDENT)
    out = orc.resample_scalar(a, g, g2, [], 2, 0)
    mid = (a[:, :, :-1].astype(np.float64) + a[:, :, 1:]) / 2
    assert np.array_equal(out[:, :, :-1], np.trunc(mid).astype(np.int16))


def test_dvf_transform_identity_outside_and_vector_resample():
    size = (9, 8, 7)
    g = orc.make_geom(size, (1, 1, 1), (0, 0, 0), IDENT)
    dvf = np.zeros((7, 8, 9, 3))
    dvf[..., 0] = 1.0  # +1 mm in x everywhere
    a = np.arange(7 * 8 * 9, dtype=np.float32).reshape(7, 8, 9)
    out = orc.resample_scalar(a, g, g, [("dvf", dvf, g)], 2, -1.0)
    assert np.allclose(out[:, :, :-1], a[:, :, 1:])
    assert np.all(out[:, :, -1] == -1.0)  # x = 8 maps to cidx 9 >= 8.5: outside the input buffer -> default
    # a field covering only part of the image: points outside the FIELD buffer are not displaced
    gs = orc.make_geom((4, 8, 7), (1, 1, 1), (0, 0, 0), IDENT)
    out2 = orc.resample_scalar(a, g, g, [("dvf", dvf[:, :, :4].copy(), gs)], 2, -1.0)
    assert np.allclose(out2[:, :, :3], a[:, :, 1:4]) and np.array_equal(out2[:, :, 4:], a[:, :, 4:])
    # vector resample of a linear field reproduces it exactly at interior sample points
    f = np.stack(np.meshgrid(np.arange(7.0), np.arange(8.0), np.arange(9.0), indexing="ij"), -1)
    g2 = orc.make_geom((5, 5, 5), (1.5, 1.25, 1.0), (0.5, 0.5, 0.5), IDENT)
    o = orc.resample_vec3(f, g, g2, [], 0.0)
    zz, yy, xx = np.meshgrid(np.arange(5), np.arange(5), np.arange(5), indexing="ij")
    assert np.allclose(o[..., 2], 0.5 + 1.5 * xx) and np.allclose(o[..., 1], 0.5 + 1.25 * yy) and np.allclose(o[..., 0], 0.5 + zz)


def test_deriche_recursive_gaussian_properties():
    # unit DC gain, symmetric impulse response close to a sampled Gaussian
    n = 64
    f = np.zeros((n, n, n, 3))
    f[n // 2, n // 2, n // 2, :] = 1.0
    g = orc.make_geom((n, n, n), (1.0, 1.0, 1.0), (0, 0, 0), IDENT)
    sigma = (2.0, 3.0, 2.5)
    out = orc.recursive_gaussian_vec3(f, g, sigma)
    assert abs(out[..., 0].sum() - 1.0) < 1e-3
    line = out[n // 2, n // 2, :, 0] / out[n // 2, n // 2, :, 0].sum()
    x = np.arange(n) - n // 2
    gauss = np.exp(-x ** 2 / (2 * sigma[0] ** 2))
    gauss /= gauss.sum()
    assert np.abs(line - gauss).max() < 5e-3
    assert np.allclose(line, line[::-1], atol=1e-12) or np.abs(line[1:] - line[1:][::-1]).max() < 1e-3
    const = np.full((8, 9, 10, 3), 3.25)
    gc = orc.make_geom((10, 9, 8), (0.9, 1.1, 2.5), (0, 0, 0), IDENT)
    assert np.abs(orc.recursive_gaussian_vec3(const, gc, (1.5, 1.5, 1.5)) - 3.25).max() < 1e-9
    with pytest.raises(RuntimeError):
        orc.recursive_gaussian_vec3(np.zeros((3, 8, 8, 3)), orc.make_geom((8, 8, 3), (1, 1, 1), (0, 0, 0), IDENT), (1, 1, 1))


def test_demons_reduces_mismatch_and_halts():
    f, m = synth_pair((40, 36, 24), seed=5, peak_mm=3.0)
    stats = []
    reg, tfm, dvf = ref.fast_symmetric_forces_demons_registration(f, m, resolution_staging=[2, 1], iteration_staging=[20, 10], level_stats=stats)
    before = np.mean((f.array - m.array) ** 2)
    after = np.mean((f.array - reg.array) ** 2)
    assert after < 0.5 * before
    assert dvf.array.shape == f.array.shape + (3,) and dvf.array.dtype == np.float64
    assert stats[0]["elapsed_iterations"] <= 20 and stats[1]["elapsed_iterations"] <= 10
    # identical images: every |F - W| < 0.001 -> zero update -> RMS 0 < 0.02 -> halts after the first iteration
    D, st = orc.demons_execute(f.array, orc.geom_of(f), f.array, orc.geom_of(f), orc.demons_params((1.5, 1.5, 1.5), 10, smooth_update_field=True))
    assert st["elapsed_iterations"] == 1 and np.all(D == 0) and st["metric"] == 0.0


def test_demons_force_sentinel_logic():
    # moving image shifted so that part of the fixed grid maps outside the moving buffer -> FLT_MAX sentinel
    f, m = synth_pair((20, 18, 16), seed=7, peak_mm=1.0)
    m2 = Image(m.array, m.GetSpacing(), (6.0, 0.0, 0.0), m.GetDirection())
    D = np.zeros(f.array.shape + (3,))
    W, U, metric, rms = orc.demons_force(f.array, orc.geom_of(f), m2.array, orc.geom_of(m2), D, orc.demons_params((1.5,) * 3, 1))
    fmax = np.finfo(np.float32).max
    assert np.all(W[:, :, :6] == fmax) and np.all(W[:, :, 6:] != fmax)
    assert np.all(U[:, :, :6, :] == 0)
    assert np.isfinite(U).all() and np.isfinite(metric) and rms > 0


def test_staple_and_fusion_tail():
    rng = np.random.default_rng(3)
    truth = np.zeros((12, 14, 16), np.uint8)
    truth[3:9, 4:11, 5:12] = 1
    raters = []
    for k in range(5):
        r = truth.copy()
        flip = rng.random(truth.shape) < 0.03 * (k + 1)
        r[flip] ^= 1
        raters.append(r)
    W, p, q, it = orc.staple(raters)
    assert 1 <= it < 100
    assert np.all(np.diff(p) < 0.02) and p[0] > 0.9 and q[0] > 0.9  # noisier raters get lower sensitivity
    assert ((W > 0.5).astype(np.uint8) == truth).mean() > 0.995
    out = ref.combine_labels_staple({str(k): {"S": Image(r)} for k, r in enumerate(raters)})["S"]
    assert out.array.dtype == np.float64 and out.array.max() == 1.0 and out.array.min() == 0.0
    assert np.all((out.array == 0) | (out.array >= 1e-4))
    # weighted vote: unweighted = mean of labels, blurred, rescaled to [0, 1]
    atlas_set = {str(k): {"DIR": {"S": Image(r), "Weight Map": Image(np.ones(r.shape, np.float32))}} for k, r in enumerate(raters)}
    prob = ref.combine_labels(atlas_set, "S")["S"].array
    assert prob.dtype == np.float32 and prob.max() == 1.0 and prob.min() == 0.0
    assert ((prob > 0.5).astype(np.uint8) == truth).mean() > 0.97


def test_reference_acceptance_sphere_phantom_dice():
    """platipy/imaging/tests/test_cardiac.py:35-71,142: sphere phantoms (radius 25 at (30+i, 64+i, 64), -1000 HU
    background, anisotropic spacing), atlas -> target Demons with staging [8,4,2] / [5,5,5] iterations,
    smoothing sigmas 0, isotropic resampling; propagated WHOLEHEART label must reach Dice > 0.99.
    (The linear pre-registration stage of the pipeline is outside this path; the atlas is placed on the
    target grid with the same geometry so Demons sees what it sees after the rigid stage.)"""
    def case(i):
        ct = insert_sphere(np.ones((60, 128, 128)) * -1000, 25, (30 + i, 64 + i, 64))
        mask = insert_sphere(np.zeros((60, 128, 128)), 25, (30 + i, 64 + i, 64))
        sp = (0.9 + 0.04 * 0.01, 0.9 + 0.04 * 0.01, 2.5 + 0.04 * 0.01)
        return Image(ct.astype(np.float32), sp, (320, -52, 60)), Image(mask.astype(np.uint8), sp, (320, -52, 60))

    target_ct, target_mask = case(4)
    atlas_ct, atlas_mask = case(2)
    _, tfm, _ = ref.fast_symmetric_forces_demons_registration(target_ct, atlas_ct, resolution_staging=[8, 4, 2], iteration_staging=[5, 5, 5],
                                                              smoothing_sigmas=[0, 0, 0], isotropic_resample=True, default_value=-1000)
    prop = ref.apply_transform(atlas_mask, target_ct, tfm, 0, sk.sitkNearestNeighbor).array
    dice0 = 2.0 * (atlas_mask.array & target_mask.array).sum() / (atlas_mask.array.sum() + target_mask.array.sum())
    dice = 2.0 * (prop & target_mask.array).sum() / (prop.sum() + target_mask.array.sum())
    assert dice > dice0
    assert dice > 0.95, (dice0, dice)


def test_oracle_is_deterministic_across_thread_counts():
    """Per-slice partial sums merged in z order: metric / RMS and the field do not depend on the OpenMP team size."""
    f, m = synth_pair((30, 26, 18), seed=9, peak_mm=2.0)
    p = orc.demons_params((1.5, 1.5, 1.5), 6, smooth_update_field=True)
    n0 = orc.num_threads()
    try:
        orc.set_num_threads(1)
        D1, s1 = orc.demons_execute(f.array, orc.geom_of(f), m.array, orc.geom_of(m), p)
        orc.set_num_threads(max(2, n0))
        D2, s2 = orc.demons_execute(f.array, orc.geom_of(f), m.array, orc.geom_of(m), p)
    finally:
        orc.set_num_threads(n0)
    assert np.array_equal(D1, D2) and s1 == s2
