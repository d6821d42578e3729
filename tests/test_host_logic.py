"""Host-side logic that needs no GPU: the SimpleITK stand-ins, transform ordering, argument validation of
the reference-facing functions, and the host parts of the library (Gaussian operator coefficients)."""
import ctypes as C

import numpy as np
import pytest

from oracle import itk_oracle as orc
from platipy_b200 import _abi
from platipy_b200 import sitk_compat as sk
from platipy_b200.sitk_compat import Image


def test_image_standin_geometry_and_cast():
    a = np.arange(24, dtype=np.float32).reshape(2, 3, 4) - 10.5
    im = Image(a, (0.9, 0.9, 2.5), (320, -52, 60))
    assert im.GetSize() == (4, 3, 2) and im.GetDepth() == 2 and im.GetPixelID() == sk.sitkFloat32
    assert im.GetDirection() == (1, 0, 0, 0, 1, 0, 0, 0, 1)
    c = sk.Cast(im, sk.sitkInt16)
    assert c.GetPixelID() == sk.sitkInt16 and c.GetSpacing() == im.GetSpacing()
    assert np.array_equal(c.array, np.trunc(a).astype(np.int16))  # static_cast truncates toward zero
    other = Image(np.zeros((2, 3, 4), np.uint8))
    other.CopyInformation(im)
    assert other.GetOrigin() == (320.0, -52.0, 60.0)
    with pytest.raises(RuntimeError):
        Image(np.zeros((3, 3, 3), np.uint8)).CopyInformation(im)
    v = sk.GetImageFromArray(np.zeros((2, 3, 4, 3)), isVector=True)
    assert v.GetPixelID() == sk.sitkVectorFloat64 and v.GetNumberOfComponentsPerPixel() == 3
    assert sk.GetArrayFromImage(im) is not im.array and np.array_equal(sk.GetArrayFromImage(im), a)
    assert bool(im) is True


def test_composite_transform_applies_last_added_first():
    t1 = sk.AffineTransform(np.eye(3) * 2.0, (0, 0, 0))
    t2 = sk.AffineTransform(np.eye(3), (1, 0, 0))
    comp = sk.CompositeTransform([t1, t2])  # p -> t1(t2(p))
    flat = comp.flatten()
    assert flat[0] is t2 and flat[1] is t1
    p = np.array([1.0, 1.0, 1.0])
    for t in flat:
        p = np.array(t.TransformPoint(p))
    assert np.allclose(p, [4.0, 2.0, 2.0])
    # centre handling of MatrixOffsetTransformBase
    t = sk.AffineTransform(np.diag([2.0, 1.0, 1.0]), (0, 0, 0), center=(3.0, 0, 0))
    assert np.allclose(t.TransformPoint((3.0, 1.0, 1.0)), (3.0, 1.0, 1.0))
    with pytest.raises(RuntimeError):
        sk.DisplacementFieldTransform(Image(np.zeros((2, 2, 2, 3), np.float32), is_vector=True))


def test_library_gaussian_operator_is_bit_identical_to_the_oracle(built):
    lib = _abi.load()
    buf = (C.c_double * 257)()
    for var, err, width in [(1.0, 0.1, 30), (2.25, 0.1, 30), (1.5 ** 2 / 0.97 ** 2, 0.1, 30), (16.0, 0.01, 128), (64.0, 0.01, 512), (0.25, 0.01, 32),
                            (64.0, 0.01, 4), (9.0, 0.1, 30)]:
        r = lib.b200reg_gaussian_operator(var, err, width, buf, 257)
        k = orc.gaussian_operator(var, err, width)
        assert 2 * r + 1 == len(k)
        assert np.array_equal(np.array(buf[: 2 * r + 1]), k)


def test_reference_facing_argument_errors(built):
    from platipy_b200 import registration as reg

    # utils.py:231-235 raises AttributeError when both factors are given -- checked before any device work
    import torch

    if torch.cuda.is_available():
        pytest.skip("argument checks that must not need a GPU are exercised on the CPU-only box")
    im = Image(np.zeros((8, 8, 8), np.float32))
    with pytest.raises(RuntimeError):  # no GPU -> loud failure, never a CPU fallback
        reg.smooth_and_resample(im, shrink_factor=2)
    with pytest.raises(RuntimeError):
        reg.fast_symmetric_forces_demons_registration(im, im)
    assert reg._check_interp(sk.sitkBSpline) == 3
    with pytest.raises(NotImplementedError):
        reg._check_interp(4)
    f = reg.FastSymmetricForcesDemonsRegistrationFilter()
    f.SetStandardDeviations(2.0)
    assert f.GetStandardDeviations() == (2.0, 2.0, 2.0)
    p = f.params(7)
    assert p.number_of_iterations == 7 and p.max_kernel_width == 30 and p.max_error == 0.1 and p.smooth_update_field == 0


def test_generate_random_augmentation_draws_like_the_reference(monkeypatch):
    """generation/augment.py:86-141: one augmentation per mask, parameters drawn from the reference's ranges; the bone mask is
    computed from the CT for the contract / expand augmentations (stubbed here: no GPU)."""
    import random

    import numpy as np

    from platipy_b200 import generation as gen
    from platipy_b200.sitk_compat import Image

    sentinel = object()
    monkeypatch.setattr(gen, "get_bone_mask", lambda image: sentinel)
    random.seed(3)
    masks = [Image(np.zeros((4, 4, 4), np.uint8), (1.0, 2.0, 2.5)) for _ in range(12)]
    ct = Image(np.zeros((4, 4, 4), np.float32))
    augs = gen.generate_random_augmentation(ct, list(masks))
    assert len(augs) == 12 and {type(a) for a in augs} <= {gen.ShiftAugment, gen.ContractAugment, gen.ExpandAugment}
    assert len({type(a) for a in augs}) >= 2
    for a in augs:
        assert isinstance(a, gen.DeformableAugment) and 3 <= a.gaussian_smooth <= 5
        if isinstance(a, gen.ShiftAugment):
            assert -10 <= a.vector_shift[0] <= 10 and a.vector_shift[1] == 10 and -10 <= a.vector_shift[2] <= 10  # (10, 10): the reference's range
        else:
            assert a.bone_mask is sentinel
        if isinstance(a, gen.ContractAugment):  # augment.py:193: millimetres -> negative voxel counts by the mask's spacing
            assert all(c <= 0 for c in a.contract) and all(isinstance(c, int) for c in a.contract)
        if isinstance(a, gen.ExpandAugment):
            assert all(0 <= v <= 10 for v in a.vector_expand)


def test_small_registration_helpers(capsys):
    import numpy as np

    from platipy_b200 import registration as reg
    from platipy_b200.sitk_compat import Image

    f = reg.FastSymmetricForcesDemonsRegistrationFilter()
    f._stats = {"elapsed_iterations": 7, "metric": 12.3456789, "rms_change": 0.1}
    reg.deformable_registration_command_iteration(f)
    assert capsys.readouterr().out == "  7 =   12.34568\n"
    img = Image(np.zeros((10, 20, 40), np.float32), (1.0, 2.0, 2.5))
    assert list(reg.control_point_spacing_distance_to_number(img, (10.0, 10.0, 10.0))) == [4, 4, 3]
