"""The named ITK-semantics switches (confidence M): every switch exists under the same name in the CPU oracle
(oracle/itk_oracle.c) and in libb200reg.so (include/b200reg.h), defaults agree, and -- CPU part -- every switch is live in the
oracle: flipping it changes the result of the operation it governs and nothing else."""
import ctypes as C

import numpy as np
import pytest

from oracle import itk_oracle as orc
from oracle import platipy_ref as ref
from platipy_b200 import _abi
from platipy_b200 import sitk_compat as sk
from platipy_b200.sitk_compat import Image
from platipy_b200.synth import smooth_random_dvf, synth_pair


def test_both_sides_declare_the_same_switches_with_the_same_defaults(built):
    lib = _abi.load()
    sem = orc.semantics()
    assert sorted(sem) == sorted(_abi.SEMANTIC_SWITCHES)
    for name, (value, meaning) in sem.items():
        assert lib.b200reg_get_semantic(name.encode()) == value, name
        assert meaning
    assert lib.b200reg_get_semantic(b"no_such_switch") == -1
    with pytest.raises(ValueError):
        _abi.set_semantic("no_such_switch", 1)
    with pytest.raises(ValueError):
        orc.set_semantic("no_such_switch", 1)
    # set / get round trip on the library side (no GPU needed: the table is host state), restored afterwards
    _abi.set_semantic("discrete_gaussian_axis_order", 1)
    assert _abi.get_semantic("discrete_gaussian_axis_order") == 1
    _abi.set_semantic("discrete_gaussian_axis_order", 0)


def _cases():
    size, sp = (20, 18, 14), (1.0, 1.2, 1.7)
    fixed, moving = synth_pair(size, seed=4, spacing=sp, peak_mm=2.0)
    g = orc.geom_of(fixed)
    g2 = orc.make_geom((17, 15, 11), (1.15, 1.4, 2.1), (0.3, 0.2, 0.1), (1, 0, 0, 0, 1, 0, 0, 0, 1))
    dvf = smooth_random_dvf(size, seed=3, peak_mm=3.0)
    ang = 0.05
    aff = sk.AffineTransform([[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1]], (0.7, -0.4, 0.3), (9.0, 9.0, 9.0))
    lab16 = Image((np.random.default_rng(0).random(size[::-1]) * 3).astype(np.int16), sp)
    moving64 = Image(moving.array.astype(np.float64), sp)
    return {
        "discrete_gaussian_axis_order": lambda: orc.discrete_gaussian_f32(fixed.array, g, [4.0, 4.0, 4.0], 32, 0.01, True),
        "recursive_gaussian_axis_order": lambda: orc.recursive_gaussian_vec3(dvf, g, [1.5, 1.5, 1.5]),
        "resample_linear_scanline": lambda: ref.apply_transform(moving64, fixed, aff, -1000, sk.sitkLinear).array,  # Float64: ulps of the index show
        # output grids that differ from the field's own: on its own grid a field is only ever read at its nodes
        "dvf_transform_interpolation": lambda: orc.resample_scalar(moving64.array, g, g2, [("dvf", dvf, g)], 2, -1000.0),
        "vector_resample_interpolation": lambda: orc.resample_vec3(dvf, g, g2, [], 0.0),
        "binary_threshold_in_pixel_type": lambda: ref._binary_threshold(lab16.array, 0.5, 255),
    }


def test_every_switch_is_live_in_the_oracle_and_independent(built):
    cases = _cases()
    base = {k: np.array(f(), copy=True) for k, f in cases.items()}
    for name in cases:
        default = orc.get_semantic(name)
        with orc.semantic(name, 1 - default):
            flipped = {k: np.array(f(), copy=True) for k, f in cases.items()}
        assert orc.get_semantic(name) == default
        for k in cases:
            same = np.array_equal(flipped[k], base[k])
            if k == name:
                assert not same, f"switch {name} does not change the operation it governs"
                # ... and only in the last bits / at ties: these are alternative roundings of the same mathematics
                if flipped[k].dtype.kind == "f":
                    assert np.allclose(flipped[k], base[k], rtol=1e-4, atol=1e-3), name
            elif not (name == "resample_linear_scanline" and k == "vector_resample_interpolation"):  # an identity resample is a linear one
                assert same, f"switch {name} changed {k}"


@pytest.mark.gpu
def test_gpu_follows_every_switch(engine):
    """Flip each switch on both sides: the CUDA path keeps matching the oracle bit for bit (so a correction found against a real
    SimpleITK is a flag flip, not an edit of CUDA code)."""
    from platipy_b200 import fusion
    from platipy_b200 import registration as reg

    size, sp = (36, 30, 22), (1.0, 1.2, 1.7)
    fixed, moving = synth_pair(size, seed=4, spacing=sp, peak_mm=2.0)
    dvf = Image(smooth_random_dvf(size, seed=3, peak_mm=3.0), sp, is_vector=True)
    tfm = sk.DisplacementFieldTransform(dvf)
    ang = 0.05
    aff = sk.AffineTransform([[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1]], (0.7, -0.4, 0.3), (9.0, 9.0, 9.0))
    lab16 = {a: {"S": Image((np.random.default_rng(a).random(size[::-1]) * 3).astype(np.int16), sp)} for a in range(3)}
    kw = dict(resolution_staging=[2, 1], iteration_staging=[6, 3])
    moving64 = Image(moving.array.astype(np.float64), sp)
    other = Image(np.zeros((15, 21, 25), np.float32), (1.15, 1.4, 2.1), (0.3, 0.2, 0.1))  # a grid that is not the field's own

    def both():
        out = {}
        out["registration"] = (reg.fast_symmetric_forces_demons_registration(fixed, moving, **kw)[2].array,
                               ref.fast_symmetric_forces_demons_registration(fixed, moving, **kw)[2].array)
        out["affine"] = (reg.apply_transform(moving, fixed, aff, -1000, sk.sitkLinear).array, ref.apply_transform(moving, fixed, aff, -1000, sk.sitkLinear).array)
        out["dvf"] = (reg.apply_transform(moving, fixed, tfm, -1000, sk.sitkLinear).array, ref.apply_transform(moving, fixed, tfm, -1000, sk.sitkLinear).array)
        out["dvf_offgrid"] = (reg.apply_transform(moving64, other, tfm, -1000, sk.sitkLinear).array, ref.apply_transform(moving64, other, tfm, -1000, sk.sitkLinear).array)
        out["smooth"] = (reg.smooth_and_resample(fixed, shrink_factor=2, smoothing_sigma=2.0).array, ref.smooth_and_resample(fixed, shrink_factor=2, smoothing_sigma=2.0).array)
        got, exp = fusion.combine_labels_staple(lab16), ref.combine_labels_staple(lab16)
        out["staple"] = (got["S"].array, exp["S"].array)
        return out

    base = both()
    for k, (g, e) in base.items():
        assert np.allclose(g, e, rtol=1e-9, atol=1e-12) if k == "staple" else np.array_equal(g, e), k
    for name in _abi.SEMANTIC_SWITCHES:
        default = orc.get_semantic(name)
        _abi.set_semantic(name, 1 - default)
        try:
            with orc.semantic(name, 1 - default):
                flipped = both()
        finally:
            _abi.set_semantic(name, default)
        changed = []
        for k, (g, e) in flipped.items():
            assert np.allclose(g, e, rtol=1e-9, atol=1e-12) if k == "staple" else np.array_equal(g, e), (name, k)
            if not np.array_equal(e, base[k][1]):
                changed.append(k)
        assert changed, f"flipping {name} changed nothing on this path"
