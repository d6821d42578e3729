"""The host-side proof behind the library's copy / per-index shortcuts (resample.cuh: identity_resample_is_exact, exported as the
diagnostic b200reg_identity_resample_is_exact): whenever it answers "exact", the resampler's arithmetic -- restated by the oracle, which
the CUDA kernels match bit for bit -- must return the input voxels themselves, for the linear and the nearest-neighbour interpolator, in
the scan-line form and in the per-voxel form.  "Not exact" is always safe, but grids that are plainly identical must be recognised,
otherwise the shortcut never triggers.  Runs on the CPU: the proof is host code."""
import ctypes as C

import numpy as np
import pytest

from oracle import itk_oracle as orc
from oracle import platipy_ref as ref
from platipy_b200 import _abi
from platipy_b200 import sitk_compat as sk
from platipy_b200.sitk_compat import Image

IDENT = (1, 0, 0, 0, 1, 0, 0, 0, 1)


def exact(a, b, scanline):
    ga, gb = _abi.geom_of(a), _abi.geom_of(b)
    return _abi.load().b200reg_identity_resample_is_exact(C.byref(ga), C.byref(gb), int(scanline))


def grids(rng, n):
    spacings = [1.0, 0.9765625, 0.7, 1.0 / 3.0, 2.5, 0.9, 1.171875, 3.0, 0.48828125]
    for _ in range(n):
        # the scan-line form divides by the row length: exact mostly for powers of two, so both kinds of sizes are drawn
        size = [int(v) for v in rng.choice([4, 8, 16, 32, 5, 12, 20, 23], size=3)]
        if rng.random() < 0.5:  # dyadic spacings and origins: the arithmetic is exact, the shortcut must trigger
            sp = [float(rng.choice([1.0, 0.5, 2.0, 0.9765625, 2.5])) for _ in range(3)]
            og = [float(rng.choice([0.0, -250.0, 64.0, -0.5])) for _ in range(3)]
        else:
            sp = [float(rng.choice(spacings)) for _ in range(3)]
            og = [float(v) for v in rng.choice([0.0, -249.51171875, 17.3, -0.1, 1e3 + 1.0 / 7.0, -33.3333], size=3)]
        yield size, sp, og


def test_identical_medical_grids_are_recognised():
    img = Image(np.zeros((100, 64, 64), np.float32), (0.9765625, 0.9765625, 2.5), (-249.51171875, -249.51171875, -120.0), IDENT)
    assert exact(img, img, 1) == 1 and exact(img, img, 0) == 1
    unit = Image(np.zeros((7, 9, 11), np.float32), (1.0, 1.0, 1.0), (0.0, 0.0, 0.0), IDENT)
    assert exact(unit, unit, 1) == 1 and exact(unit, unit, 0) == 1


def test_different_grids_are_rejected():
    a = Image(np.zeros((8, 9, 10), np.float32), (1.0, 1.0, 2.0), (0.0, 0.0, 0.0), IDENT)
    assert exact(a, Image(np.zeros((8, 9, 11), np.float32), (1.0, 1.0, 2.0), (0.0, 0.0, 0.0), IDENT), 1) == 0           # size
    assert exact(a, Image(np.zeros((8, 9, 10), np.float32), (1.0, 1.0, 2.0), (0.5, 0.0, 0.0), IDENT), 1) == 0           # half a voxel off
    assert exact(a, Image(np.zeros((8, 9, 10), np.float32), (1.0, 1.0, 2.0), (1e-9, 0.0, 0.0), IDENT), 1) == 0          # a nanometre off
    assert exact(a, Image(np.zeros((8, 9, 10), np.float32), (1.0, 1.0000001, 2.0), (0.0, 0.0, 0.0), IDENT), 1) == 0     # spacing
    c, s = np.cos(0.01), np.sin(0.01)
    rot = (c, -s, 0, s, c, 0, 0, 0, 1)
    r = Image(np.zeros((8, 9, 10), np.float32), (1.0, 1.0, 2.0), (0.0, 0.0, 0.0), rot)
    assert exact(r, r, 1) == 0                                                                                             # oriented grids take the kernel
    flip = (-1, 0, 0, 0, -1, 0, 0, 0, 1)
    f = Image(np.zeros((8, 9, 10), np.float32), (1.0, 1.0, 2.0), (0.0, 0.0, 0.0), flip)
    assert exact(f, f, 1) == 0


@pytest.mark.parametrize("scanline", [1, 0])
def test_exact_means_the_resampler_returns_the_input(scanline):
    rng = np.random.default_rng(7 + scanline)
    n_exact = 0
    with orc.semantic("resample_linear_scanline", scanline):
        _abi.set_semantic("resample_linear_scanline", scanline)
        try:
            for size, sp, og in grids(rng, 60):
                arr = (rng.normal(size=size[::-1]) * 1000).astype(np.float32)
                img = Image(arr, sp, og, IDENT)
                # the same grid, and a grid whose origin was rebuilt by arithmetic that is only ALMOST the same
                og2 = [o + s * 3 - s * 3 for o, s in zip(og, sp)]
                for other in (img, Image(arr, sp, og2, IDENT)):
                    if exact(img, other, scanline):
                        n_exact += 1
                        for interp in (sk.sitkLinear, sk.sitkNearestNeighbor):
                            out = ref.resample(img, other, None, interp, -7.0)
                            assert np.array_equal(out.array, arr), (size, sp, og, interp)
        finally:
            _abi.set_semantic("resample_linear_scanline", 1)
    print("exact:", n_exact, "of 120")
    assert n_exact >= 25  # the proof is not vacuous on ordinary grids
