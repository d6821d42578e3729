"""GPU parity of the binary post-processing (platipy/imaging/label/fusion.py:295-328 process_probability_image:
BinaryFillhole, ConnectedComponent + largest object) against the CPU oracle -- integer work, bit-exact."""
import numpy as np
import pytest
import scipy.ndimage as ndi

from oracle import itk_oracle as orc
from oracle import platipy_ref as ref
from platipy_b200 import fusion
from platipy_b200.sitk_compat import Image

pytestmark = pytest.mark.gpu


def _blobs(shape, seed, thr=0.55, sigma=2.0):
    rng = np.random.default_rng(seed)
    v = ndi.gaussian_filter(rng.standard_normal(shape), sigma)
    v = (v - v.min()) / (v.max() - v.min())
    return (v > thr).astype(np.uint8)


SHAPES = [(20, 33, 47), (9, 64, 70), (1, 40, 40), (40, 1, 33), (17, 31, 1), (30, 30, 30), (5, 7, 100)]


def test_fillhole_bit_exact(engine):
    for seed, shape in enumerate(SHAPES):
        for thr in (0.45, 0.6):
            m = _blobs(shape, seed, thr)
            got = engine.to_host(engine.binary_fillhole(engine.to_device(Image(m)))).array
            assert np.array_equal(got, orc.binary_fillhole(m)), (shape, thr)
    # noise: many tiny holes and objects, runs crossing the 32-voxel chunks
    m = (np.random.default_rng(9).random((12, 50, 130)) > 0.4).astype(np.uint8)
    got = engine.to_host(engine.binary_fillhole(engine.to_device(Image(m)))).array
    assert np.array_equal(got, orc.binary_fillhole(m))


def test_largest_component_bit_exact(engine):
    for seed, shape in enumerate(SHAPES):
        for thr in (0.5, 0.62):
            m = _blobs(shape, 20 + seed, thr)
            d, info = engine.largest_component(engine.to_device(Image(m)), want_info=True)
            exp, ncomp, nvox = orc.largest_component(m)
            assert np.array_equal(engine.to_host(d).array, exp), (shape, thr)
            assert info["n_components"] == ncomp and info["voxels"] == nvox
    # ties: equal-size objects, the first in raster order wins
    m = np.zeros((4, 10, 70), np.uint8)
    m[3, 9, 60:66] = 1
    m[0, 2, 30:36] = 1
    m[2, 5, 0:6] = 1
    d, info = engine.largest_component(engine.to_device(Image(m)), want_info=True)
    exp, ncomp, nvox = orc.largest_component(m)
    assert np.array_equal(engine.to_host(d).array, exp) and info["n_components"] == 3 and info["voxels"] == 6
    # percolating noise: one huge ragged object
    m = (np.random.default_rng(3).random((16, 40, 90)) > 0.55).astype(np.uint8)
    d = engine.largest_component(engine.to_device(Image(m)))
    assert np.array_equal(engine.to_host(d).array, orc.largest_component(m)[0])
    # empty and full
    for m in (np.zeros((3, 5, 40), np.uint8), np.ones((3, 5, 40), np.uint8)):
        d, info = engine.largest_component(engine.to_device(Image(m)), want_info=True)
        assert np.array_equal(engine.to_host(d).array, m) and info["n_components"] == int(m.any())


def test_process_probability_image_matches_oracle(engine):
    rng = np.random.default_rng(11)
    for dtype in (np.float32, np.float64):
        for seed, shape in enumerate(SHAPES[:4] + [(24, 48, 80)]):
            v = ndi.gaussian_filter(rng.standard_normal(shape), 2.5)
            p = np.clip((v - v.min()) / (v.max() - v.min()) * 0.9, 0, 1).astype(dtype)
            img = Image(p, (1.0, 1.2, 2.0), (3.0, -4.0, 5.0))
            for thr in (0.5, 0.37):
                got = fusion.process_probability_image(img, thr)
                exp = ref.process_probability_image(img, thr)
                assert got.array.dtype == np.uint8 and got.GetSpacing() == img.GetSpacing() and got.GetOrigin() == img.GetOrigin()
                assert np.array_equal(got.array, exp.array), (dtype, shape, thr)
    assert fusion.process_probability_image(np.zeros((4, 6, 8), np.float32)).array.sum() == 0
    with pytest.raises(NotImplementedError):
        fusion.process_probability_image(Image(np.ones((2, 2, 2), np.uint8)))


def test_full_size_properties(engine):
    """512 x 512 x 256: idempotence of both operators, monotonicity of fill-hole, and the largest object of the
    filled mask is a subset of it (size-independent properties; the oracle's flood fill is too slow here)."""
    import torch

    shape = (256, 512, 512)
    g = torch.Generator(device="cuda").manual_seed(7)
    v = torch.rand(shape, generator=g, device="cuda")
    # smooth a little so that objects and holes of many sizes exist
    k = torch.ones((1, 1, 3, 3, 3), device="cuda") / 27.0
    v = torch.nn.functional.conv3d(v[None, None], k, padding=1)[0, 0]
    m = (v > 0.5).to(torch.uint8)
    from platipy_b200.engine import DeviceImage

    d = DeviceImage(m.contiguous(), np.uint8, (1.0, 1.0, 1.0), (0.0, 0.0, 0.0), (1, 0, 0, 0, 1, 0, 0, 0, 1), False)
    d = engine.to_device(d)  # orders the engine's stream after the torch stream that produced the mask
    f1 = engine.binary_fillhole(d)
    f2 = engine.binary_fillhole(f1)
    engine.synchronize()
    assert torch.equal(f1.tensor, f2.tensor)
    assert bool((f1.tensor >= m).all())
    l1, info = engine.largest_component(f1, want_info=True)
    l2, info2 = engine.largest_component(l1, want_info=True)
    engine.synchronize()
    assert torch.equal(l1.tensor, l2.tensor) and info2["n_components"] == 1 and info2["voxels"] == info["voxels"]
    assert int(l1.tensor.sum().item()) == info["voxels"] and bool((l1.tensor <= f1.tensor).all())
