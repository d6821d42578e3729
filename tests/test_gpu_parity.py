"""GPU parity tests: every CUDA entry point against the CPU oracle on the same seeded inputs, through the
C ABI (ctypes -> libb200reg.so).  Integer / index work and everything computed in the oracle's operation
order is required to be BIT-EXACT; the documented floating-point tolerances are the north-star ones:
DVF within 1e-4 mm per component, resampled float intensities within 1e-5 relative, labels bit-exact."""
import numpy as np
import pytest

from oracle import itk_oracle as orc
from oracle import platipy_ref as ref
from platipy_b200 import registration as reg
from platipy_b200 import sitk_compat as sk
from platipy_b200.sitk_compat import Image
from platipy_b200.synth import smooth_random_dvf, synth_labels, synth_pair

pytestmark = pytest.mark.gpu

IDENT = (1, 0, 0, 0, 1, 0, 0, 0, 1)
DVF_TOL_MM = 1e-4
REL_TOL = 1e-5


def rot_direction(ax=0.05, az=0.1):
    cx, sx, cz, sz = np.cos(ax), np.sin(ax), np.cos(az), np.sin(az)
    rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    return tuple((rz @ rx).reshape(9))


def host(engine, d):
    return engine.to_host(d, pinned=False)


# ---------------------------------------------------------------------------------------------------------
def test_discrete_gaussian_bit_exact(engine):
    rng = np.random.default_rng(0)
    for size, spacing, var, mw in [((33, 20, 17), (1.0, 1.0, 1.0), 4.0, 32), ((24, 31, 9), (0.9, 0.9, 2.5), (16.0, 16.0, 16.0), 128),
                                   ((16, 16, 16), (1.0, 1.0, 1.0), 64.0, 3), ((70, 5, 6), (0.5, 1.0, 1.0), (1.0, 0.0001, 9.0), 32)]:
        a = (rng.normal(size=size[::-1]) * 300).astype(np.float32)
        im = Image(a, spacing)
        out = host(engine, engine.discrete_gaussian(engine.to_device(im), var, mw)).array
        exp = orc.discrete_gaussian_f32(a, orc.geom_of(im), var, mw)
        assert np.array_equal(out, exp)


@pytest.mark.parametrize("dtype", [np.uint8, np.int8, np.int16, np.uint16, np.int32, np.uint32, np.int64, np.uint64, np.float32, np.float64])
def test_resample_all_pixel_types_bit_exact(engine, dtype):
    rng = np.random.default_rng(1)
    if np.issubdtype(dtype, np.integer):
        info = np.iinfo(dtype)
        a = rng.integers(max(info.min, -30000), min(info.max, 30000), size=(11, 13, 15)).astype(dtype)
    else:
        a = (rng.normal(size=(11, 13, 15)) * 500).astype(dtype)
    src = Image(a, (0.9, 1.1, 2.5), (3.0, -2.0, 10.0), rot_direction())
    ref_img = Image(np.zeros((9, 17, 12), np.uint8), (1.3, 0.8, 2.0), (2.0, -1.0, 9.0), rot_direction(0.02, -0.07))
    th = 0.08
    aff = sk.AffineTransform([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1.02]], (0.6, -0.3, 0.4), center=(8, 6, 20))
    dvf = Image(smooth_random_dvf((10, 9, 8), seed=3, peak_mm=2.0), (1.7, 1.9, 3.0), (1.0, -3.0, 8.0), IDENT, True)
    dtf = sk.DisplacementFieldTransform(dvf)
    for tfm in [None, aff, dtf, sk.CompositeTransform([aff, dtf]), sk.CompositeTransform([dtf, aff])]:
        for interp in (sk.sitkNearestNeighbor, sk.sitkLinear):
            out = reg.apply_transform(src, ref_img, tfm, default_value=-7 if np.issubdtype(dtype, np.signedinteger) or np.issubdtype(dtype, np.floating) else 3,
                                      interpolator=interp)
            exp = ref.apply_transform(src, ref_img, tfm, default_value=-7 if np.issubdtype(dtype, np.signedinteger) or np.issubdtype(dtype, np.floating) else 3,
                                      interpolator=interp)
            assert out.array.dtype == dtype and out.GetSize() == ref_img.GetSize() and out.GetSpacing() == ref_img.GetSpacing()
            assert np.array_equal(out.array, exp.array), (tfm, interp)
    # reference_image=None -> the input's own grid (utils.py:178-181)
    out = reg.apply_transform(src, None, aff, 0, sk.sitkLinear)
    assert np.array_equal(out.array, ref.apply_transform(src, None, aff, 0, sk.sitkLinear).array)


def test_resample_batch_equals_single_calls(engine):
    fixed, moving = synth_pair((40, 36, 28), seed=2, spacing=(1.0, 1.0, 1.5))
    labels = [Image(l, fixed.GetSpacing()) for l in synth_labels((40, 36, 28), 11, seed=200)]
    dvf = Image(smooth_random_dvf((40, 36, 28), seed=5, peak_mm=4.0), fixed.GetSpacing(), fixed.GetOrigin(), IDENT, True)
    tfm = sk.DisplacementFieldTransform(dvf)
    imgs = [moving] + labels
    outs = reg.apply_transform_batch(imgs, fixed, tfm, [-1000] + [0] * 11, [sk.sitkLinear] + [sk.sitkNearestNeighbor] * 11)
    for im, o, dv, ip in zip(imgs, outs, [-1000] + [0] * 11, [sk.sitkLinear] + [sk.sitkNearestNeighbor] * 11):
        exp = ref.apply_transform(im, fixed, tfm, dv, ip)
        assert np.array_equal(o.array, exp.array)
    with pytest.raises(NotImplementedError):
        reg.apply_transform(moving, fixed, tfm, 0, 4)  # sitkGaussian: not one of the interpolators the path implements


@pytest.mark.parametrize("moving_direction", [IDENT, "rotated"])
def test_float32_resample_through_field_on_output_grid(engine, moving_direction):
    """One Float32 image, linear interpolation, one displacement field that lives on the output grid (deformable.py:139-140, :281-301): this
    call takes the Demons loop's warp kernel (demons_split.cuh: resample_f32_on_grid_dvf).  Same bits as the oracle's generic resampler,
    including points pushed outside the moving buffer (default value) and an oriented moving image (the kernel's general-geometry form)."""
    fixed, moving = synth_pair((45, 38, 27), seed=61, spacing=(1.1, 0.9, 2.0), origin=(4.0, -3.0, 12.0))
    direction = rot_direction(0.04, -0.06) if moving_direction == "rotated" else IDENT
    moving = Image(moving.array[:25, :30, :41].copy(), (1.0, 1.2, 1.9), (5.0, -1.0, 13.0), direction)  # its own grid, smaller than the fixed one
    dvf = Image(smooth_random_dvf((45, 38, 27), seed=62, peak_mm=14.0), fixed.GetSpacing(), fixed.GetOrigin(), IDENT, True)
    tfm = sk.DisplacementFieldTransform(dvf)
    for dv in (-1000, 0, 3.5e38, 1e39):  # the last one is beyond Float32: clamped like ITK's CastPixelWithBoundsChecking
        got = reg.apply_transform(moving, fixed, tfm, dv, sk.sitkLinear)
        exp = ref.apply_transform(moving, fixed, tfm, dv, sk.sitkLinear)
        assert got.array.dtype == np.float32
        assert np.array_equal(got.array, exp.array), dv
        assert (got.array == np.float32(min(dv, float(np.finfo(np.float32).max)))).any()  # some points do leave the moving buffer
    # one label at a time, nearest neighbour (the per-structure calls of multiatlas/run.py:338-345): resample_nn_on_grid_dvf, raw-bit transport
    rng = np.random.default_rng(64)
    for dtype, dv in ((np.uint8, 0), (np.uint8, 300), (np.int8, -7), (np.int16, -1000), (np.uint16, 70000), (np.int32, -5), (np.uint32, 7), (np.float32, -1e39)):
        if np.issubdtype(dtype, np.integer):
            info = np.iinfo(dtype)
            arr = rng.integers(max(info.min, -30000), min(info.max, 30000), size=moving.array.shape).astype(dtype)
        else:
            arr = (rng.normal(size=moving.array.shape) * 500).astype(dtype)
        lab = Image(arr, moving.GetSpacing(), moving.GetOrigin(), moving.GetDirection())
        got = reg.apply_transform(lab, fixed, tfm, dv, sk.sitkNearestNeighbor)
        exp = ref.apply_transform(lab, fixed, tfm, dv, sk.sitkNearestNeighbor)
        assert got.array.dtype == dtype
        assert np.array_equal(got.array, exp.array), (dtype, dv)
    # on the moving image's own grid (the level-start warp: output grid = input grid = field grid)
    dvf_m = Image(smooth_random_dvf((41, 30, 25), seed=63, peak_mm=6.0), moving.GetSpacing(), moving.GetOrigin(), IDENT, True)
    if moving_direction == IDENT:
        t2 = sk.DisplacementFieldTransform(dvf_m)
        assert np.array_equal(reg.apply_transform(moving, moving, t2, 0, sk.sitkLinear).array, ref.apply_transform(moving, moving, t2, 0, sk.sitkLinear).array)


def test_bspline_interpolation_bit_exact(engine):
    """sitk.sitkBSpline (order 3; deformable.py:221-224, utils.py:148-192): coefficient decomposition + 64-point
    evaluation are the same IEEE operations as the oracle's restatement of itk::BSplineInterpolateImageFunction."""
    fixed, moving = synth_pair((44, 37, 23), seed=21, spacing=(1.0, 1.3, 2.1), origin=(-3.0, 8.0, 1.5))
    dvf = Image(smooth_random_dvf((44, 37, 23), seed=22, peak_mm=5.0), fixed.GetSpacing(), fixed.GetOrigin(), fixed.GetDirection(), is_vector=True)
    tfm = sk.DisplacementFieldTransform(dvf)
    aff = sk.AffineTransform([[0.99, -0.05, 0.01], [0.05, 1.01, 0.0], [0.0, 0.02, 0.98]], (1.5, -2.0, 0.7), (20.0, 20.0, 20.0))
    small = Image(moving.array[:3, :9, :5].copy(), moving.GetSpacing())  # lines shorter than the filter horizon
    cases = [(moving, fixed, tfm, -1000), (moving, fixed, aff, -1000), (moving, None, None, 0), (small, small, sk.AffineTransform(translation=(0.4, 0.3, 0.2)), 0),
             (Image(moving.array.astype(np.int16), moving.GetSpacing(), moving.GetOrigin()), fixed, aff, -1000),
             (Image((moving.array > -500).astype(np.uint8) * 200, moving.GetSpacing(), moving.GetOrigin()), fixed, tfm, 0),
             (Image(moving.array.astype(np.float64), moving.GetSpacing(), moving.GetOrigin()), fixed, tfm, 0)]
    for img, ref_img, t, dv in cases:
        got = reg.apply_transform(img, ref_img, t, dv, sk.sitkBSpline)
        exp = ref.apply_transform(img, ref_img, t, dv, sk.sitkBSpline)
        assert got.array.dtype == img.array.dtype
        assert np.array_equal(got.array, exp.array), (img.array.dtype, type(t).__name__)
    # the Demons driver accepts interp_order=3 as well (pyramid, per-level and final resampling, deformable.py:281-304)
    kw = dict(resolution_staging=[2, 1], iteration_staging=[5, 3], interp_order=sk.sitkBSpline)
    img, _, dvf_g = reg.fast_symmetric_forces_demons_registration(fixed, moving, **kw)
    img_o, _, dvf_o = ref.fast_symmetric_forces_demons_registration(fixed, moving, **kw)
    assert np.abs(dvf_g.array - dvf_o.array).max() <= DVF_TOL_MM
    assert np.abs(img.array - img_o.array).max() <= REL_TOL * max(1.0, np.abs(img_o.array).max())


def test_resample_vec3_and_compose_bit_exact(engine):
    f = Image(smooth_random_dvf((20, 18, 16), seed=7, peak_mm=3.0), (2.0, 2.1, 3.0), (5, 5, 5), rot_direction(), True)
    grid = Image(np.zeros((31, 35, 39), np.uint8), (1.0, 1.05, 1.5), (5.2, 4.9, 5.1), rot_direction())
    d = engine.to_device(f)
    out = host(engine, engine.resample_vec3(d, grid)).array
    exp = ref.resample(f, grid).array
    assert np.array_equal(out, exp)
    # dvf_total + Resample(dvf_iter, DisplacementFieldTransform(dvf_total))  (deformable.py:154)
    tot = Image(smooth_random_dvf((20, 18, 16), seed=8, peak_mm=5.0), f.GetSpacing(), f.GetOrigin(), f.GetDirection(), True)
    tfm = sk.DisplacementFieldTransform(sk.Cast(tot, sk.sitkVectorFloat64))
    exp2 = tot.array + ref.resample(f, f, tfm).array
    got = host(engine, engine.compose_dvf(engine.to_device(tot), d)).array
    assert np.array_equal(got, exp2)


def test_recursive_gaussian_bit_exact(engine):
    arr = smooth_random_dvf((23, 19, 14), seed=9, peak_mm=3.0) + np.random.default_rng(0).normal(size=(14, 19, 23, 3))
    f = Image(arr, (0.9, 0.9, 2.5), is_vector=True)
    sigma = (1.5 / 0.9, 1.5 / 0.9, 1.5 / 2.5)  # the reference passes voxel-unit numbers as physical sigmas
    got = host(engine, engine.recursive_gaussian(engine.to_device(f), sigma)).array
    exp = orc.recursive_gaussian_vec3(f.array, orc.geom_of(f), sigma)
    assert np.array_equal(got, exp)
    with pytest.raises(RuntimeError):
        engine.recursive_gaussian(engine.to_device(Image(np.zeros((3, 8, 8, 3)), is_vector=True)), (1, 1, 1))


def _params(std, iters, smooth_update=True):
    f = reg.FastSymmetricForcesDemonsRegistrationFilter()
    f.SetStandardDeviations(std)
    f.SetSmoothUpdateField(smooth_update)
    return f.params(iters)


def test_demons_force_and_smoothing_bit_exact(engine):
    fixed, moving = synth_pair((37, 29, 21), seed=11, spacing=(0.97, 0.97, 2.0), origin=(-20.0, 4.0, 100.0), peak_mm=3.0)
    # moving image on a different, shifted grid so part of the fixed grid maps outside it (FLT_MAX logic)
    moving = Image(moving.array[:, :, 3:], moving.GetSpacing(), (-20.0 + 5 * 0.97, 4.0, 100.0), IDENT)
    D = smooth_random_dvf((37, 29, 21), seed=12, peak_mm=2.0)
    std = (1.5 / 0.97, 1.5 / 0.97, 0.75)
    W, U, metric, rms = orc.demons_force(fixed.array, orc.geom_of(fixed), moving.array, orc.geom_of(moving), D, orc.demons_params(std, 1))
    dF, dM = engine.to_device(fixed), engine.to_device(moving)
    dD = engine.to_device(Image(D, fixed.GetSpacing(), fixed.GetOrigin(), IDENT, True))
    gW, gU, gmetric, grms = engine.demons_force(dF, dM, dD, _params(std, 1))
    assert (W == np.finfo(np.float32).max).sum() > 100
    assert np.array_equal(host(engine, gW).array, W)
    assert np.array_equal(host(engine, gU).array, U)
    assert abs(gmetric - metric) <= 1e-12 * abs(metric) and abs(grms - rms) <= 1e-12 * abs(rms)
    # PDE smoothing (x, y, z; clamp boundary) of a field
    sm = orc.pde_smooth_field(U, orc.geom_of(fixed), std)
    gsm = host(engine, engine.pde_smooth_field(gU, std)).array
    assert np.array_equal(gsm, sm)


def test_demons_execute_config1(engine):
    """BASELINE.json configs[0]: 64x64x32, one level, 10 iterations."""
    fixed, moving = synth_pair((64, 64, 32), seed=0)
    std = (1.5, 1.5, 1.5)
    D, st = orc.demons_execute(fixed.array, orc.geom_of(fixed), moving.array, orc.geom_of(moving), orc.demons_params(std, 10, smooth_update_field=True))
    gD, gst = engine.demons_execute(engine.to_device(fixed), engine.to_device(moving), _params(std, 10))
    got = host(engine, gD).array
    assert gst["elapsed_iterations"] == st["elapsed_iterations"] == 10
    assert np.abs(got - D).max() <= DVF_TOL_MM
    assert np.array_equal(got, D), "same operation order as the oracle: expected bit-exact"
    assert abs(gst["metric"] - st["metric"]) <= 1e-10 * st["metric"] and abs(gst["rms_change"] - st["rms_change"]) <= 1e-10


def test_demons_halt_rules(engine):
    fixed, _ = synth_pair((32, 32, 16), seed=1)
    dF = engine.to_device(fixed)
    # identical images -> zero update -> RMS 0 < 0.02 -> exactly one iteration, zero field
    gD, st = engine.demons_execute(dF, dF, _params((1.5,) * 3, 25))
    assert st["elapsed_iterations"] == 1 and st["metric"] == 0.0 and float(gD.tensor.abs().max()) == 0.0
    # zero iterations -> zero field, nothing elapsed
    gD, st = engine.demons_execute(dF, dF, _params((1.5,) * 3, 0))
    assert st["elapsed_iterations"] == 0 and float(gD.tensor.abs().max()) == 0.0
    # early stop happens at the same iteration as in the oracle
    f2, m2 = synth_pair((32, 32, 16), seed=4, peak_mm=0.3, noise_hu=0.0)
    p = orc.demons_params((1.5,) * 3, 200, smooth_update_field=True)
    D, so = orc.demons_execute(f2.array, orc.geom_of(f2), m2.array, orc.geom_of(m2), p)
    gD, sg = engine.demons_execute(engine.to_device(f2), engine.to_device(m2), _params((1.5,) * 3, 200))
    assert sg["elapsed_iterations"] == so["elapsed_iterations"]
    assert np.abs(host(engine, gD).array - D).max() <= DVF_TOL_MM


CASES = {
    "cfg1_sigma0": dict(size=(64, 64, 32), kw=dict(resolution_staging=[1], iteration_staging=[10], smoothing_sigmas=[0])),
    "cfg1_default_sigma": dict(size=(64, 64, 32), kw=dict(resolution_staging=[1], iteration_staging=[10])),
    "three_levels": dict(size=(64, 56, 40), kw=dict(resolution_staging=[4, 2, 1], iteration_staging=[20, 10, 5])),
    "platipy_defaults": dict(size=(72, 64, 48), kw=dict()),
    "anisotropic_offset_rotated": dict(size=(60, 50, 28), spacing=(0.9, 0.9, 2.5), origin=(320.0, -52.0, 60.0), direction=rot_direction(0.03, 0.05),
                                       kw=dict(resolution_staging=[2, 1], iteration_staging=[10, 5], regularisation_kernel_mm=[1.5, 1.5, 2.0])),
    "isotropic_resample": dict(size=(60, 50, 28), spacing=(0.9, 0.9, 2.5), origin=(320.0, -52.0, 60.0),
                               kw=dict(resolution_staging=[6, 3, 1.5], iteration_staging=[10, 8, 5], isotropic_resample=True, smoothing_sigmas=[0, 0, 0])),
    # regularisation radius 8 voxels (> the fused kernel's limit): separable large-radius smoothing path
    "fine_spacing_large_radius": dict(size=(56, 48, 40), spacing=(0.3, 0.3, 0.3), kw=dict(resolution_staging=[2, 1], iteration_staging=[6, 4])),
    # different x / y radii (3 vs 2) and z radius 1: generic fused kernel path
    "anisotropic_radii": dict(size=(56, 48, 24), spacing=(0.75, 1.0, 2.5), kw=dict(resolution_staging=[2, 1], iteration_staging=[6, 4])),
    # very small volume (tiles mostly empty) and a non-smoothed field (update smoothing only is fixed on by the reference)
    "tiny_volume": dict(size=(9, 7, 6), kw=dict(resolution_staging=[1], iteration_staging=[5], smoothing_sigmas=[0])),
    "nn_interp_int16": dict(size=(48, 40, 24), dtype=np.int16, kw=dict(resolution_staging=[2, 1], iteration_staging=[6, 4], interp_order=sk.sitkNearestNeighbor)),
}


@pytest.mark.parametrize("name", list(CASES))
def test_fast_symmetric_forces_demons_registration_parity(engine, name):
    c = CASES[name]
    fixed, moving = synth_pair(c["size"], seed=21, spacing=c.get("spacing", (1.0, 1.0, 1.0)), origin=c.get("origin", (0.0, 0.0, 0.0)), peak_mm=4.0)
    if "direction" in c:
        fixed = Image(fixed.array, fixed.GetSpacing(), fixed.GetOrigin(), c["direction"])
        moving = Image(moving.array, moving.GetSpacing(), moving.GetOrigin(), c["direction"])
    if "dtype" in c:
        fixed, moving = sk.Cast(fixed, sk.dtype_to_pixel_id(c["dtype"])), sk.Cast(moving, sk.dtype_to_pixel_id(c["dtype"]))
    img, tfm, dvf = reg.fast_symmetric_forces_demons_registration(fixed, moving, **c["kw"])
    stats = []
    img_o, tfm_o, dvf_o = ref.fast_symmetric_forces_demons_registration(fixed, moving, level_stats=stats, **c["kw"])
    assert dvf.GetPixelID() == sk.sitkVectorFloat64 and dvf.GetSize() == fixed.GetSize() and dvf.GetSpacing() == fixed.GetSpacing()
    assert img.GetPixelID() == moving.GetPixelID() and img.GetOrigin() == fixed.GetOrigin()
    err = np.abs(dvf.array - dvf_o.array).max()
    assert err <= DVF_TOL_MM, f"DVF max error {err} mm"
    if np.issubdtype(img.array.dtype, np.floating):
        assert np.abs(img.array - img_o.array).max() <= REL_TOL * max(1.0, np.abs(img_o.array).max())
    else:
        assert np.array_equal(img.array, img_o.array)
    # the transform object is usable by apply_transform (labels bit-exact)
    lab = Image(synth_labels(c["size"], 1, seed=300)[0], fixed.GetSpacing(), fixed.GetOrigin(), fixed.GetDirection())
    assert np.array_equal(reg.apply_transform(lab, fixed, tfm, 0, sk.sitkNearestNeighbor).array,
                          ref.apply_transform(lab, fixed, tfm_o, 0, sk.sitkNearestNeighbor).array)


def test_initial_displacement_field_and_device_io(engine):
    fixed, moving = synth_pair((40, 40, 24), seed=31, peak_mm=3.0)
    init = Image(smooth_random_dvf((20, 20, 12), seed=32, peak_mm=1.0), (2.0, 2.0, 2.0), (0.5, 0.5, 0.5), IDENT, True)
    kw = dict(resolution_staging=[2, 1], iteration_staging=[5, 5], initial_displacement_field=init)
    _, _, dvf = reg.fast_symmetric_forces_demons_registration(fixed, moving, **kw)
    _, _, dvf_o = ref.fast_symmetric_forces_demons_registration(fixed, moving, **kw)
    assert np.abs(dvf.array - dvf_o.array).max() <= DVF_TOL_MM
    # device in -> device out, nothing copied back
    dF, dM = engine.to_device(fixed), engine.to_device(moving)
    img_d, tfm_d, dvf_d = reg.fast_symmetric_forces_demons_registration(dF, dM, resolution_staging=[2, 1], iteration_staging=[5, 5])
    from platipy_b200.engine import DeviceImage

    assert isinstance(img_d, DeviceImage) and isinstance(dvf_d, DeviceImage)
    _, _, dvf_h = reg.fast_symmetric_forces_demons_registration(fixed, moving, resolution_staging=[2, 1], iteration_staging=[5, 5])
    assert np.array_equal(host(engine, dvf_d).array, dvf_h.array)
    lab = engine.to_device(Image(synth_labels((40, 40, 24), 1)[0]))
    out = reg.apply_transform(lab, dF, tfm_d, 0, sk.sitkNearestNeighbor)
    assert isinstance(out, DeviceImage)


def test_multiscale_demons_with_initial_transform(engine):
    """deformable.py:101-108: the initial field sampled from a transform (TransformToDisplacementField)."""
    fixed, moving = synth_pair((40, 36, 24), seed=61, spacing=(1.0, 1.1, 1.5), peak_mm=3.0)
    th = 0.03
    aff = sk.AffineTransform([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1]], (0.8, -0.5, 0.3), center=(20, 20, 18))
    kw = dict(resolution_staging=[2, 1], smoothing_sigmas=[2, 1], iteration_staging=[5, 5], initial_transform=aff)
    f_reg = reg.FastSymmetricForcesDemonsRegistrationFilter()
    f_reg.SetSmoothUpdateField(True)
    f_reg.SetStandardDeviations((1.5, 1.5 / 1.1, 1.0))
    f_ref = ref.DemonsFilter()
    f_ref.SetSmoothUpdateField(True)
    f_ref.SetStandardDeviations((1.5, 1.5 / 1.1, 1.0))
    got = reg.multiscale_demons(f_reg, fixed, moving, **kw)
    exp = ref.multiscale_demons(f_ref, fixed, moving, **kw)
    assert np.abs(got.array - exp.array).max() <= DVF_TOL_MM
    # the sampled field itself
    d = engine.to_host(engine.transform_to_dvf(aff, engine.to_device(fixed)), pinned=False).array
    assert np.array_equal(d, orc.transform_to_dvf(orc.geom_of(fixed), [("affine", aff.matrix, aff.offset)]))


def test_smooth_and_resample_parity_and_errors(engine):
    fixed, _ = synth_pair((50, 44, 30), seed=41, spacing=(0.9, 0.9, 2.5))
    for kw in [dict(shrink_factor=2, smoothing_sigma=2), dict(shrink_factor=[2, 2, 1], smoothing_sigma=[1.0, 1.0, 2.5]), dict(isotropic_voxel_size_mm=3, smoothing_sigma=0),
               dict(smoothing_sigma=1.5), dict(shrink_factor=4, interpolator=sk.sitkNearestNeighbor)]:
        out, exp = reg.smooth_and_resample(fixed, **kw), ref.smooth_and_resample(fixed, **kw)
        assert out.GetSize() == exp.GetSize() and np.allclose(out.GetSpacing(), exp.GetSpacing(), rtol=0, atol=0)
        assert np.array_equal(out.array, exp.array), kw
    with pytest.raises(AttributeError):
        reg.smooth_and_resample(fixed, isotropic_voxel_size_mm=2, shrink_factor=2)
    with pytest.raises(ZeroDivisionError):
        reg.smooth_and_resample(fixed, shrink_factor=64)  # a level collapsing to one voxel: utils.py:252-255 divides by zero


@pytest.mark.parametrize("size,spacing,kw", [
    ((96, 80, 48), (1.0, 1.0, 1.0), dict(shrink_factor=4, smoothing_sigma=2.0)),
    ((97, 83, 41), (0.9, 1.1, 2.5), dict(shrink_factor=8, smoothing_sigma=4.0)),
    ((64, 64, 40), (0.98, 0.98, 3.0), dict(isotropic_voxel_size_mm=6, smoothing_sigma=3.0)),      # z hardly shrinks: that axis stays (almost) full
    ((75, 50, 33), (1.0, 1.0, 1.0), dict(shrink_factor=3, smoothing_sigma=1.5)),                 # odd factor: continuous indices on / near integers
    ((80, 72, 36), (1.2, 1.2, 1.2), dict(shrink_factor=[4, 4, 2], smoothing_sigma=[2.4, 2.4, 1.2])),
    ((130, 24, 20), (1.0, 1.0, 1.0), dict(shrink_factor=[16, 2, 2], smoothing_sigma=[6.0, 1.0, 1.0])),  # radius near the supported maximum
])
def test_restricted_pyramid_level_is_bit_identical(engine, size, spacing, kw):
    """A shrinking smooth_and_resample blurs only the planes / rows / columns the level's linear interpolation reads (pyramid.cuh): equal, bit
    for bit, to blur-everything-then-resample (the generic path, forced through allow_restricted=0) and to the oracle's restatement of
    utils.py:195-267 -- with both settings of the two semantic switches the restricted path depends on."""
    from platipy_b200 import _abi

    img, _ = synth_pair(size, seed=sum(size), spacing=spacing, origin=(-12.5, 7.25, 100.0))
    exp = ref.smooth_and_resample(img, **kw)
    out = reg.smooth_and_resample(img, **kw)
    assert out.GetSize() == exp.GetSize()
    assert np.array_equal(out.array, exp.array)
    # the generic path on the same grid
    d = engine.to_device(img)
    sig = kw["smoothing_sigma"]
    var = [s * s for s in sig] if hasattr(sig, "__iter__") else [sig * sig] * 3
    mw = int(max(8 * v * sp for sp, v in zip(spacing, var)))
    generic = host(engine, engine.smooth_and_resample(d, var, mw, exp, sk.sitkLinear, allow_restricted=0)).array
    assert np.array_equal(generic, exp.array)
    forced = host(engine, engine.smooth_and_resample(d, var, mw, exp, sk.sitkLinear, allow_restricted=2)).array  # also below the pay-off threshold
    assert np.array_equal(forced, exp.array)
    for name in ("discrete_gaussian_axis_order", "resample_linear_scanline"):
        flipped = 1 - _abi.get_semantic(name)
        _abi.set_semantic(name, flipped)
        try:
            with orc.semantic(name, flipped):
                exp2 = ref.smooth_and_resample(img, **kw).array
            got2 = host(engine, engine.smooth_and_resample(d, var, mw, exp, sk.sitkLinear, allow_restricted=2)).array
        finally:
            _abi.set_semantic(name, 1 - flipped)
        assert np.array_equal(got2, exp2), name


def test_reference_acceptance_sphere_phantom_dice_gpu(engine):
    """The reference's only acceptance criterion for this path (test_cardiac.py:35-71,142), on the GPU."""
    from platipy_b200.synth import insert_sphere

    def case(i):
        ct = insert_sphere(np.ones((60, 128, 128)) * -1000, 25, (30 + i, 64 + i, 64))
        mask = insert_sphere(np.zeros((60, 128, 128)), 25, (30 + i, 64 + i, 64))
        sp = (0.94, 0.94, 2.54)
        return Image(ct.astype(np.float32), sp, (320, -52, 60)), Image(mask.astype(np.uint8), sp, (320, -52, 60))

    target_ct, target_mask = case(4)
    atlas_ct, atlas_mask = case(2)
    kw = dict(resolution_staging=[8, 4, 2], iteration_staging=[5, 5, 5], smoothing_sigmas=[0, 0, 0], isotropic_resample=True, default_value=-1000)
    _, tfm, dvf = reg.fast_symmetric_forces_demons_registration(target_ct, atlas_ct, **kw)
    _, tfm_o, dvf_o = ref.fast_symmetric_forces_demons_registration(target_ct, atlas_ct, **kw)
    assert np.abs(dvf.array - dvf_o.array).max() <= DVF_TOL_MM
    prop = reg.apply_transform(atlas_mask, target_ct, tfm, 0, sk.sitkNearestNeighbor).array
    assert np.array_equal(prop, ref.apply_transform(atlas_mask, target_ct, tfm_o, 0, sk.sitkNearestNeighbor).array)
    dice = 2.0 * (prop & target_mask.array).sum() / (prop.sum() + target_mask.array.sum())
    assert dice > 0.95


def test_randomised_resample_chains(engine):
    """40 random (geometry, transform chain, pixel type, interpolator) draws: bit-exact against the oracle."""
    rng = np.random.default_rng(2024)
    dtypes = [np.uint8, np.int16, np.uint16, np.int32, np.float32, np.float64]
    for trial in range(40):
        sz_in = tuple(int(v) for v in rng.integers(5, 24, 3))
        sz_out = tuple(int(v) for v in rng.integers(4, 26, 3))
        dt = dtypes[trial % len(dtypes)]
        a = (rng.normal(size=sz_in[::-1]) * 200).astype(dt) if np.issubdtype(dt, np.floating) else rng.integers(0, 200, sz_in[::-1]).astype(dt)
        src = Image(a, tuple(rng.uniform(0.5, 2.5, 3)), tuple(rng.uniform(-20, 20, 3)), rot_direction(*rng.uniform(-0.2, 0.2, 2)))
        grid = Image(np.zeros(sz_out[::-1], np.uint8), tuple(rng.uniform(0.5, 2.5, 3)), tuple(np.array(src.GetOrigin()) + rng.uniform(-3, 3, 3)),
                     rot_direction(*rng.uniform(-0.2, 0.2, 2)))
        parts = []
        for _ in range(int(rng.integers(0, 4))):
            if rng.random() < 0.5:
                m = np.eye(3) + rng.normal(scale=0.05, size=(3, 3))
                parts.append(sk.AffineTransform(m, rng.uniform(-2, 2, 3), center=rng.uniform(0, 10, 3)))
            else:
                fs = tuple(int(v) for v in rng.integers(4, 12, 3))
                f = Image(rng.normal(scale=1.5, size=fs[::-1] + (3,)), tuple(rng.uniform(1.0, 4.0, 3)), tuple(np.array(src.GetOrigin()) + rng.uniform(-5, 5, 3)),
                          IDENT, True)
                parts.append(sk.DisplacementFieldTransform(f))
        tfm = sk.CompositeTransform(parts) if parts else None
        interp = sk.sitkLinear if trial % 2 else sk.sitkNearestNeighbor
        dv = float(rng.integers(0, 50))
        got = reg.apply_transform(src, grid, tfm, dv, interp)
        exp = ref.apply_transform(src, grid, tfm, dv, interp)
        assert np.array_equal(got.array, exp.array), trial


def test_randomised_small_registrations(engine):
    """Random small volumes / spacings / stagings: DVF within 1e-4 mm (in fact identical) of the oracle."""
    rng = np.random.default_rng(7)
    for trial in range(6):
        size = tuple(int(v) for v in rng.integers(14, 40, 3))
        sp = tuple(float(v) for v in rng.uniform(0.6, 2.6, 3))
        fixed, moving = synth_pair(size, seed=100 + trial, spacing=sp, origin=tuple(rng.uniform(-50, 50, 3)), peak_mm=2.0)
        kw = dict(resolution_staging=[2, 1], iteration_staging=[int(rng.integers(1, 7)), int(rng.integers(1, 5))],
                  regularisation_kernel_mm=float(rng.uniform(1.0, 2.5)), smoothing_sigma_factor=float(rng.uniform(0.5, 1.5)))
        _, _, dvf = reg.fast_symmetric_forces_demons_registration(fixed, moving, **kw)
        _, _, dvf_o = ref.fast_symmetric_forces_demons_registration(fixed, moving, **kw)
        assert np.abs(dvf.array - dvf_o.array).max() <= DVF_TOL_MM, (trial, size, sp)


def test_iteration_events_and_trace_match_the_oracle(engine, capsys):
    """IterationEvent callbacks (deformable.py:260-264, utils.py:37-41): GetElapsedIterations() / GetMetric() at every iteration equal
    the oracle's per-iteration metric trace; ``verbose=True`` prints the reference's line once per iteration of every level."""
    fixed, moving = synth_pair((40, 36, 24), seed=11, spacing=(1.0, 1.0, 1.5))
    flt = reg.FastSymmetricForcesDemonsRegistrationFilter()
    flt.SetStandardDeviations((1.5, 1.5, 1.0))
    flt.SetSmoothUpdateField(True)
    flt.SetNumberOfIterations(12)
    seen = []
    flt.AddCommand(sk.sitkIterationEvent, lambda: seen.append((flt.GetElapsedIterations(), flt.GetMetric(), flt.GetRMSChange())))
    flt.AddCommand(sk.sitkStartEvent, lambda: seen.append("start event callbacks are not iteration callbacks"))
    flt.Execute(fixed, moving)
    _, st = orc.demons_execute(fixed.array, orc.geom_of(fixed), moving.array, orc.geom_of(moving),
                               orc.demons_params((1.5, 1.5, 1.0), 12, smooth_update_field=True), trace=True)
    assert [s[0] for s in seen] == list(range(1, st["elapsed_iterations"] + 1))
    assert np.allclose([s[1] for s in seen], st["metric_trace"], rtol=1e-10, atol=0)
    assert flt.GetElapsedIterations() == st["elapsed_iterations"] and abs(flt.GetMetric() - st["metric"]) <= 1e-10 * st["metric"]
    assert abs(seen[-1][2] - st["rms_change"]) <= 1e-10 * max(st["rms_change"], 1e-300)
    # verbose: one "{elapsed:3} = {metric:10.5f}" line per iteration and level
    capsys.readouterr()
    reg.fast_symmetric_forces_demons_registration(fixed, moving, resolution_staging=[2, 1], iteration_staging=[5, 3], verbose=True)
    lines = [l for l in capsys.readouterr().out.splitlines() if " = " in l]
    elapsed = [s["elapsed_iterations"] for s in reg.LAST_LEVEL_STATS]
    assert len(lines) == sum(elapsed)
    assert [int(l.split("=")[0]) for l in lines] == [i + 1 for n in elapsed for i in range(n)]


def test_pipelined_host_api_gives_the_same_results(engine):
    """submit_registration / iter_registrations / register_batch overlap the PCIe copies of back-to-back registrations; the values are
    those of the synchronous call, bit for bit, and the host results live in pinned memory."""
    kw = dict(resolution_staging=[2, 1], iteration_staging=[6, 4])
    pairs = [synth_pair((40, 36, 24), seed=20 + k, spacing=(1.0, 1.0, 1.5)) for k in range(3)]
    pairs[1] = (Image(pairs[1][0].array.astype(np.int16), (1.0, 1.0, 1.5)), Image(pairs[1][1].array.astype(np.int16), (1.0, 1.0, 1.5)))
    sync = [reg.fast_symmetric_forces_demons_registration(f, m, **kw) for f, m in pairs]
    batch = reg.register_batch(pairs, **kw)
    assert len(batch) == 3
    for (i0, t0, d0), (i1, t1, d1) in zip(sync, batch):
        assert np.array_equal(d0.array, d1.array) and np.array_equal(i0.array, i1.array) and i0.array.dtype == i1.array.dtype
        assert np.array_equal(t1.GetDisplacementField().array, d1.array)
    pend = reg.submit_registration(*pairs[0], **kw)
    img, tfm, dvf = pend.result()
    assert pend.done() and np.array_equal(dvf.array, sync[0][2].array) and [s["elapsed_iterations"] for s in pend.level_stats] == [6, 4]
    # the transform that came back is usable as it is
    assert np.array_equal(reg.apply_transform(pairs[0][1], pairs[0][0], tfm, -1000, sk.sitkLinear).array, img.array)
    assert list(reg.iter_registrations([], **kw)) == []


@pytest.mark.parametrize("size", [(40, 36, 24), (37, 33, 21)])
def test_bit_packed_label_propagation_is_exact(engine, size):
    """Batches with many UInt8 nearest-neighbour items take the bit-packed path (one gathered word per output voxel instead of one
    byte per structure): same bits as the per-image calls and as the oracle -- for {0, 1} and {0, 255} masks, an empty mask, a
    multi-valued label (recognised on the device, gathered from its own image), non-zero default values, more items than one word
    holds (40 -> two groups), volumes whose size is / is not a multiple of four, and next to a linearly interpolated CT."""
    sp = (1.0, 1.2, 1.6)
    fixed, moving = synth_pair(size, seed=31, spacing=sp)
    rng = np.random.default_rng(3)
    base = synth_labels(size, 40, seed=700)
    labels = []
    for k, l in enumerate(base):
        if k % 5 == 1:
            l = (l * 255).astype(np.uint8)
        elif k % 5 == 2:
            l = (l * rng.integers(1, 4, size=l.shape)).astype(np.uint8)   # several non-zero values
        elif k % 5 == 3:
            l = np.zeros_like(l)
        labels.append(Image(l, sp))
    dvf = Image(smooth_random_dvf(size, seed=5, peak_mm=6.0), sp, is_vector=True)
    tfm = sk.DisplacementFieldTransform(dvf)
    defaults = [-1000] + [0 if k % 7 else 9 for k in range(40)]
    interps = [sk.sitkLinear] + [sk.sitkNearestNeighbor] * 40
    outs = reg.apply_transform_batch([moving] + labels, fixed, tfm, defaults, interps)
    g = orc.geom_of(fixed)
    chain = [("dvf", dvf.array, g)]
    assert np.array_equal(outs[0].array, orc.resample_scalar(moving.array, g, g, chain, 2, -1000.0))
    for k in range(40):
        one = reg.apply_transform(labels[k], fixed, tfm, defaults[k + 1], sk.sitkNearestNeighbor)   # per-image path (no packing)
        assert np.array_equal(outs[k + 1].array, one.array), k
        assert np.array_equal(outs[k + 1].array, orc.resample_scalar(labels[k].array, g, g, chain, 1, defaults[k + 1])), k
    # an affine chain onto another grid (points outside the input get the default value)
    aff = sk.AffineTransform([[1, 0.02, 0], [-0.02, 1, 0], [0, 0, 1]], (3.0, -2.0, 1.0), (15.0, 15.0, 15.0))
    outs = reg.apply_transform_batch(labels[:9], fixed, aff, [5] * 9, [sk.sitkNearestNeighbor] * 9)
    for k in range(9):
        assert np.array_equal(outs[k].array, ref.apply_transform(labels[k], fixed, aff, 5, sk.sitkNearestNeighbor).array), k


def test_exponentiate_field_matches_oracle(engine):
    """Scaling and squaring (the north star's optional "exponentiation"; not on the reference path): automatic and fixed N, bit-exact
    against the oracle's composition chain; exp of a zero field is zero; the result inverts well against exp(-v)."""
    size, sp = (36, 30, 22), (1.0, 1.2, 1.7)
    v = Image(smooth_random_dvf(size, seed=13, peak_mm=5.0), sp, is_vector=True)
    for n in (None, 0, 3):
        got = reg.exponentiate_field(v, n)
        exp = ref.exponentiate_field(v, n)
        assert np.array_equal(got.array, exp.array), n
    zero = Image(np.zeros(size[::-1] + (3,)), sp, is_vector=True)
    assert not reg.exponentiate_field(zero).array.any()
    # exp(v) o exp(-v) ~ identity: compose and look at the residual away from the border
    fwd, bwd = reg.exponentiate_field(v), reg.exponentiate_field(Image(-v.array, sp, is_vector=True))
    comp = ref.resample(fwd, fwd, sk.DisplacementFieldTransform(bwd)).array + bwd.array
    assert np.abs(comp[4:-4, 4:-4, 4:-4]).max() < 0.05 * np.abs(v.array).max()


def test_fast_mode_is_an_explicit_close_but_not_bit_equal_path(engine):
    """precision="fast": float32 fields + float32 FMA smoothing inside the Demons loop.  Not a parity path -- it must be asked
    for, the default stays bit-identical to the oracle, and what it returns stays close to the parity field (the thresholds of the ESM
    force amplify float32 rounding here and there, so the bar is statistical: median and 99th percentile)."""
    fixed, moving = synth_pair((64, 48, 40), seed=17, spacing=(1.0, 1.0, 1.5), peak_mm=3.0)
    kw = dict(resolution_staging=[2, 1], iteration_staging=[12, 8])
    _, _, d_par = reg.fast_symmetric_forces_demons_registration(fixed, moving, **kw)
    st_par = [s["elapsed_iterations"] for s in reg.LAST_LEVEL_STATS]
    _, _, d_fast = reg.fast_symmetric_forces_demons_registration(fixed, moving, precision="fast", **kw)
    st_fast = [s["elapsed_iterations"] for s in reg.LAST_LEVEL_STATS]
    _, _, d_par2 = reg.fast_symmetric_forces_demons_registration(fixed, moving, precision="parity", **kw)
    assert np.array_equal(d_par.array, d_par2.array) and np.array_equal(d_par.array, ref.fast_symmetric_forces_demons_registration(fixed, moving, **kw)[2].array)
    assert st_par == st_fast == [12, 8]
    err = np.abs(d_fast.array - d_par.array)
    assert err.max() > 0.0, "fast mode returned the parity bits: the switch did not act"
    assert np.median(err) < 1e-3 and np.percentile(err, 99) < 0.1 and err.max() < 1.0, (np.median(err), np.percentile(err, 99), err.max())
    # a row length the TMA unit cannot take in float32 (not a multiple of 4) silently stays in parity mode
    f2, m2 = synth_pair((46, 40, 24), seed=18, spacing=(1.0, 1.0, 1.5), peak_mm=2.0)
    a = reg.fast_symmetric_forces_demons_registration(f2, m2, precision="fast", resolution_staging=[1], iteration_staging=[4])[2]
    b = reg.fast_symmetric_forces_demons_registration(f2, m2, resolution_staging=[1], iteration_staging=[4])[2]
    assert np.array_equal(a.array, b.array)
    with pytest.raises(ValueError):
        reg.fast_symmetric_forces_demons_registration(fixed, moving, precision="sloppy", **kw)
