"""CPU tests of the NIfTI-1 reader / writer (platipy_b200/nifti_io.py; the on-disk format of the atlas pipeline,
multiatlas/run.py:160-164): header layout, RAS <-> LPS geometry, every pixel type, gzip, rescaling, error cases."""
import gzip
import struct

import numpy as np
import pytest

from platipy_b200 import nifti_io as nio
from platipy_b200 import sitk_compat as sk
from platipy_b200.sitk_compat import Image


def _rot(axis, ang):
    axis = np.asarray(axis, float) / np.linalg.norm(axis)
    k = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + np.sin(ang) * k + (1 - np.cos(ang)) * k @ k


@pytest.mark.parametrize("dtype", [np.uint8, np.int8, np.int16, np.uint16, np.int32, np.uint32, np.int64, np.uint64, np.float32, np.float64])
@pytest.mark.parametrize("ext", [".nii", ".nii.gz"])
def test_round_trip_every_pixel_type(tmp_path, dtype, ext):
    rng = np.random.default_rng(0)
    arr = (rng.random((5, 7, 9)) * 100).astype(dtype)
    d = _rot((0.3, -0.2, 1.0), 0.4)
    img = Image(arr, (0.9765625, 1.25, 2.5), (-123.5, 47.25, 1020.0), tuple(d.reshape(9)))
    p = tmp_path / f"x{ext}"
    sk.WriteImage(img, str(p))
    back = sk.ReadImage(str(p))
    assert back.array.dtype == arr.dtype and np.array_equal(back.array, arr)
    assert back.GetSize() == img.GetSize()
    assert np.allclose(back.GetSpacing(), img.GetSpacing(), rtol=1e-6)
    assert np.allclose(back.GetOrigin(), img.GetOrigin(), rtol=1e-6)
    assert np.allclose(back.GetDirection(), img.GetDirection(), atol=2e-6)  # header geometry is float32


def test_header_fields_and_lps_ras_convention(tmp_path):
    img = Image(np.arange(24, dtype=np.int16).reshape(2, 3, 4), (1.0, 2.0, 3.0), (10.0, 20.0, 30.0))
    p = tmp_path / "h.nii"
    nio.write_image(img, p)
    raw = p.read_bytes()
    assert len(raw) == 352 + 24 * 2 and struct.unpack("<i", raw[:4])[0] == 348 and raw[344:348] == b"n+1\0"
    assert struct.unpack("<8h", raw[40:56])[:4] == (3, 4, 3, 2)
    assert struct.unpack("<hh", raw[70:74]) == (4, 16)
    assert struct.unpack("<f", raw[108:112])[0] == 352.0
    assert struct.unpack("<hh", raw[252:256]) == (1, 1)
    srow = np.array(struct.unpack("<12f", raw[280:328])).reshape(3, 4)
    # identity LPS direction is diag(-1, -1, 1) in RAS; the origin's x and y change sign
    assert np.allclose(srow, [[-1, 0, 0, -10], [0, -2, 0, -20], [0, 0, 3, 30]])
    assert np.allclose(struct.unpack("<3f", raw[268:280]), [-10, -20, 30])
    # quaternion of diag(-1, -1, 1): 180 degrees about z -> (b, c, d) = (0, 0, 1)
    assert np.allclose(struct.unpack("<3f", raw[256:268]), [0, 0, 1])


def _hand_header(dim, datatype, bitpix, pixdim, qform_code=0, sform_code=0, quat=(0, 0, 0, 0, 0, 0), srow=None, slope=0.0, inter=0.0, endian="<"):
    h = bytearray(352)
    struct.pack_into(endian + "i", h, 0, 348)
    struct.pack_into(endian + "8h", h, 40, *dim)
    struct.pack_into(endian + "hh", h, 70, datatype, bitpix)
    struct.pack_into(endian + "8f", h, 76, *pixdim)
    struct.pack_into(endian + "fff", h, 108, 352.0, slope, inter)
    struct.pack_into(endian + "hh", h, 252, qform_code, sform_code)
    struct.pack_into(endian + "6f", h, 256, *quat)
    if srow is not None:
        struct.pack_into(endian + "12f", h, 280, *np.asarray(srow, float).reshape(12))
    h[344:348] = b"n+1\0"
    return bytes(h)


def test_reads_sform_only_big_endian_and_rescaled_files(tmp_path):
    data = np.arange(24, dtype=">i2").reshape(2, 3, 4)
    # RAS sform: axes swapped and scaled; no qform
    srow = [[0, -2.0, 0, 5.0], [1.5, 0, 0, -7.0], [0, 0, 4.0, 11.0]]
    p = tmp_path / "s.nii"
    p.write_bytes(_hand_header((3, 4, 3, 2, 1, 1, 1, 1), 4, 16, (1, 1.5, 2.0, 4.0, 0, 0, 0, 0), 0, 2, srow=srow, endian=">") + data.tobytes())
    img = nio.read_image(p)
    assert np.array_equal(img.array, data.astype(np.int16)) and img.GetSpacing() == (1.5, 2.0, 4.0)
    assert np.allclose(img.GetOrigin(), (-5.0, 7.0, 11.0))
    assert np.allclose(np.array(img.GetDirection()).reshape(3, 3), [[0, 1, 0], [-1, 0, 0], [0, 0, 1]])
    # no transform at all: identity geometry with pixdim spacing; slope / intercept rescale to float32
    q = tmp_path / "r.nii.gz"
    with gzip.open(q, "wb") as f:
        f.write(_hand_header((3, 4, 3, 2, 1, 1, 1, 1), 2, 8, (1, 0.5, 0.5, 2.0, 0, 0, 0, 0), slope=2.0, inter=-1000.0) + np.arange(24, dtype=np.uint8).tobytes())
    img = nio.read_image(q)
    assert img.array.dtype == np.float32 and np.array_equal(img.array.ravel(), np.arange(24, dtype=np.float32) * 2 - 1000)
    assert img.GetSpacing() == (0.5, 0.5, 2.0) and img.GetOrigin() == (0.0, 0.0, 0.0) and img.GetDirection()[0] == 1.0
    # qform with a negative qfac (left-handed index space)
    r = tmp_path / "q.nii"
    r.write_bytes(_hand_header((3, 2, 2, 2, 1, 1, 1, 1), 16, 32, (-1, 1, 1, 1, 0, 0, 0, 0), 1, 0, quat=(0, 0, 0, 1, 2, 3)) + np.zeros(8, np.float32).tobytes())
    img = nio.read_image(r)
    assert np.allclose(np.array(img.GetDirection()).reshape(3, 3), np.diag([-1.0, -1.0, -1.0])) and np.allclose(img.GetOrigin(), (-1, -2, 3))


def test_quaternion_round_trip_and_errors(tmp_path):
    rng = np.random.default_rng(1)
    for _ in range(50):
        r = _rot(rng.standard_normal(3), rng.uniform(-np.pi, np.pi))
        if rng.random() < 0.3:
            r[:, 2] = -r[:, 2]
        b, c, d, qfac = nio.matrix_to_quaternion(r)
        assert np.allclose(nio.quaternion_to_matrix(b, c, d, qfac), r, atol=1e-12)
    bad = tmp_path / "bad.nii"
    bad.write_bytes(b"\0" * 400)
    with pytest.raises(RuntimeError):
        nio.read_image(bad)
    with pytest.raises(NotImplementedError):
        nio.write_image(Image(np.zeros((2, 2, 2, 3), np.complex64), is_vector=True), tmp_path / "v.nii")
    four_d = tmp_path / "4d.nii"
    four_d.write_bytes(_hand_header((4, 2, 2, 2, 5, 1, 1, 1), 2, 8, (1, 1, 1, 1, 1, 0, 0, 0)) + bytes(40))
    with pytest.raises(NotImplementedError):
        nio.read_image(four_d)


def test_vector_image_round_trip(tmp_path):
    """Displacement fields (sitkVectorFloat64 / Float32) as itk::NiftiImageIO lays them out: dim = (5, x, y, z, 1, 3),
    NIFTI_INTENT_VECTOR, components planar in the file and interleaved in memory, values stored as they are."""
    import struct

    rng = np.random.default_rng(3)
    r = _rot(rng.standard_normal(3), 0.4)
    for dtype, name in ((np.float64, "dvf.nii.gz"), (np.float32, "dvf32.nii")):
        field = Image(rng.standard_normal((5, 6, 7, 3)).astype(dtype), (0.9, 1.1, 2.5), (-12.0, 30.5, 4.0), tuple(r.reshape(9)), is_vector=True)
        path = tmp_path / name
        sk.WriteImage(field, path)
        back = sk.ReadImage(path)
        assert back.is_vector and back.array.dtype == dtype and back.array.shape == (5, 6, 7, 3)
        assert np.array_equal(back.array, field.array)
        assert back.GetPixelID() == (sk.sitkVectorFloat64 if dtype == np.float64 else sk.sitkVectorFloat32)
        assert np.allclose(back.GetSpacing(), field.GetSpacing(), rtol=1e-6) and np.allclose(back.GetOrigin(), field.GetOrigin(), rtol=1e-6)
        assert np.allclose(back.GetDirection(), field.GetDirection(), atol=2e-6)  # header geometry is float32
    raw = (tmp_path / "dvf32.nii").read_bytes()
    assert struct.unpack("<8h", raw[40:56])[:6] == (5, 7, 6, 5, 1, 3) and struct.unpack("<h", raw[68:70])[0] == 1007
    planes = np.frombuffer(raw, np.float32, offset=352).reshape(3, 5, 6, 7)
    assert np.array_equal(planes[1], field.array[..., 1])  # component 1 is one contiguous block of the file
    tfm = sk.DisplacementFieldTransform(sk.ReadImage(tmp_path / "dvf.nii.gz"))  # what a caller does with a stored field
    assert tfm.GetDisplacementField().array.shape == (5, 6, 7, 3)
