"""
NIfTI-1 reading and writing for 3-D scalar images and 3-D vector images (displacement fields) -- the on-disk format
either side of the atlas pipeline (reference multiatlas/run.py:160-164 ``sitk.ReadImage``; service code writes results
and deformation fields with ``sitk.WriteImage``).

Host-side only (numpy + gzip); no device work.  Geometry follows itk::NiftiImageIO:

* NIfTI stores an index -> RAS+ millimetre transform, ITK / SimpleITK images live in LPS+: the first two rows of the
  transform change sign in both directions.
* Reading uses the qform (quaternion + pixdim + qoffset, ``qfac`` = pixdim[0] flips the third axis) when
  ``qform_code > 0``, otherwise the sform rows when ``sform_code > 0``, otherwise spacing from pixdim with an identity
  direction and zero origin.  Spacing is always pixdim[1..3]; with the sform the direction columns are the normalised
  matrix columns.
* Writing sets both transforms (codes NIFTI_XFORM_SCANNER_ANAT = 1), ``xyzt_units`` = mm | s, ``scl_slope`` = 1,
  ``scl_inter`` = 0, single-file magic ``n+1`` with ``vox_offset`` = 352.  Header geometry fields are float32, so a
  round trip reproduces spacing / origin / direction to float32 precision (as with SimpleITK).
* ``scl_slope`` / ``scl_inter`` other than (0 | 1, 0) rescale the data to float32 on reading, like ITK.

* Vector images follow itk::NiftiImageIO: ``dim`` = (5, nx, ny, nz, 1, components), ``intent_code`` =
  NIFTI_INTENT_VECTOR (1007), components stored as the slowest file dimension (planar) and interleaved in memory;
  component values are stored as they are (ITK's ConvertRASVectors is off by default).

Paths ending in ``.gz`` are gzip-compressed.
"""
from __future__ import annotations

import gzip
import struct

import numpy as np

from .sitk_compat import Image

# NIfTI datatype code <-> numpy dtype
_CODE_TO_DTYPE = {2: np.uint8, 4: np.int16, 8: np.int32, 16: np.float32, 64: np.float64, 256: np.int8, 512: np.uint16, 768: np.uint32,
                  1024: np.int64, 1280: np.uint64}
_DTYPE_TO_CODE = {np.dtype(v): k for k, v in _CODE_TO_DTYPE.items()}


def _open(path, mode):
    return gzip.open(path, mode) if str(path).endswith(".gz") else open(path, mode)


def quaternion_to_matrix(b, c, d, qfac):
    a2 = 1.0 - (b * b + c * c + d * d)
    if a2 < 1e-7:  # nifti_quatern_to_mat44: special case of a 180 degree rotation
        a = 0.0
        n = 1.0 / np.sqrt(b * b + c * c + d * d)
        b, c, d = b * n, c * n, d * n
    else:
        a = np.sqrt(a2)
    r = np.array([[a * a + b * b - c * c - d * d, 2 * b * c - 2 * a * d, 2 * b * d + 2 * a * c],
                  [2 * b * c + 2 * a * d, a * a + c * c - b * b - d * d, 2 * c * d - 2 * a * b],
                  [2 * b * d - 2 * a * c, 2 * c * d + 2 * a * b, a * a + d * d - c * c - b * b]])
    if qfac < 0:
        r[:, 2] = -r[:, 2]
    return r


def matrix_to_quaternion(r):
    """nifti_mat44_to_quatern for a proper or improper rotation matrix: returns (b, c, d, qfac)."""
    r = np.array(r, dtype=np.float64)
    qfac = 1.0
    if np.linalg.det(r) < 0:
        r[:, 2] = -r[:, 2]
        qfac = -1.0
    a = r[0, 0] + r[1, 1] + r[2, 2] + 1.0
    if a > 0.5:
        a = 0.5 * np.sqrt(a)
        b = 0.25 * (r[2, 1] - r[1, 2]) / a
        c = 0.25 * (r[0, 2] - r[2, 0]) / a
        d = 0.25 * (r[1, 0] - r[0, 1]) / a
    else:
        xd, yd, zd = 1.0 + r[0, 0] - (r[1, 1] + r[2, 2]), 1.0 + r[1, 1] - (r[0, 0] + r[2, 2]), 1.0 + r[2, 2] - (r[0, 0] + r[1, 1])
        if xd > 1.0:
            b = 0.5 * np.sqrt(xd)
            c = 0.25 * (r[0, 1] + r[1, 0]) / b
            d = 0.25 * (r[0, 2] + r[2, 0]) / b
            a = 0.25 * (r[2, 1] - r[1, 2]) / b
        elif yd > 1.0:
            c = 0.5 * np.sqrt(yd)
            b = 0.25 * (r[0, 1] + r[1, 0]) / c
            d = 0.25 * (r[1, 2] + r[2, 1]) / c
            a = 0.25 * (r[0, 2] - r[2, 0]) / c
        else:
            d = 0.5 * np.sqrt(zd)
            b = 0.25 * (r[0, 2] + r[2, 0]) / d
            c = 0.25 * (r[1, 2] + r[2, 1]) / d
            a = 0.25 * (r[1, 0] - r[0, 1]) / d
        if a < 0.0:
            b, c, d = -b, -c, -d
    return float(b), float(c), float(d), qfac


_LPS = np.diag([-1.0, -1.0, 1.0])  # RAS <-> LPS


def read_image(path):
    """``sitk.ReadImage`` for a 3-D scalar NIfTI-1 file."""
    with _open(path, "rb") as f:
        raw = f.read()
    if len(raw) < 348:
        raise RuntimeError(f"{path}: not a NIfTI-1 file (shorter than the header)")
    little = struct.unpack("<i", raw[:4])[0] == 348
    if not little and struct.unpack(">i", raw[:4])[0] != 348:
        raise RuntimeError(f"{path}: not a NIfTI-1 file (sizeof_hdr != 348)")
    e = "<" if little else ">"
    magic = raw[344:348]
    if magic not in (b"n+1\0", b"ni1\0"):
        raise RuntimeError(f"{path}: not a NIfTI-1 file (magic {magic!r})")
    if magic == b"ni1\0":
        raise NotImplementedError("two-file NIfTI (.hdr/.img) is not supported")
    dim = struct.unpack(e + "8h", raw[40:56])
    datatype, = struct.unpack(e + "h", raw[70:72])
    pixdim = struct.unpack(e + "8f", raw[76:108])
    vox_offset, scl_slope, scl_inter = struct.unpack(e + "fff", raw[108:120])
    qform_code, sform_code = struct.unpack(e + "hh", raw[252:256])
    qb, qc, qd, qx, qy, qz = struct.unpack(e + "6f", raw[256:280])
    srow = np.array(struct.unpack(e + "12f", raw[280:328]), dtype=np.float64).reshape(3, 4)
    ndim = dim[0]
    ncomp = int(dim[5]) if ndim == 5 and dim[4] <= 1 and dim[5] > 1 else 1
    if ndim < 1 or ndim > 7 or (ncomp == 1 and any(d > 1 for d in dim[4:ndim + 1])) or any(d > 1 for d in dim[6:ndim + 1]):
        raise NotImplementedError(f"{path}: only 3-D scalar and 3-D vector images are supported (dim = {dim})")
    nx, ny, nz = (max(int(dim[k]), 1) if k <= ndim else 1 for k in (1, 2, 3))
    if datatype not in _CODE_TO_DTYPE:
        raise NotImplementedError(f"{path}: NIfTI datatype {datatype} is not supported")
    dt = np.dtype(_CODE_TO_DTYPE[datatype]).newbyteorder(e)
    off = int(vox_offset) if vox_offset >= 352 else 352
    arr = np.frombuffer(raw, dtype=dt, count=ncomp * nx * ny * nz, offset=off).astype(dt.newbyteorder("="), copy=True)
    # the file holds one plane set per component; a vector image interleaves them: [z, y, x, c]
    arr = arr.reshape(nz, ny, nx) if ncomp == 1 else np.ascontiguousarray(np.moveaxis(arr.reshape(ncomp, nz, ny, nx), 0, -1))
    if scl_slope not in (0.0, 1.0) or (scl_slope != 0.0 and scl_inter != 0.0):
        arr = (arr.astype(np.float32) * np.float32(scl_slope) + np.float32(scl_inter)).astype(np.float32)
    spacing = np.array([abs(pixdim[k]) if k <= ndim and pixdim[k] != 0 else 1.0 for k in (1, 2, 3)], dtype=np.float64)
    if qform_code > 0:
        qfac = -1.0 if pixdim[0] < 0 else 1.0
        rot = quaternion_to_matrix(qb, qc, qd, qfac)
        direction = _LPS @ rot
        origin = _LPS @ np.array([qx, qy, qz], dtype=np.float64)
    elif sform_code > 0:
        m = srow[:, :3]
        norms = np.sqrt((m * m).sum(axis=0))
        norms[norms == 0] = 1.0
        direction = _LPS @ (m / norms)
        origin = _LPS @ srow[:, 3]
    else:
        direction, origin = np.eye(3), np.zeros(3)
    return Image(arr, tuple(spacing), tuple(origin), tuple(direction.reshape(9)), is_vector=ncomp > 1)


def write_image(image, path):
    """``sitk.WriteImage`` for a 3-D scalar or vector image as single-file NIfTI-1 (``.nii`` or ``.nii.gz``)."""
    from . import sitk_compat as sk

    image = sk.to_native(image)
    arr = np.ascontiguousarray(image.array)
    if arr.dtype == np.bool_:
        arr = arr.astype(np.uint8)
    if arr.dtype not in _DTYPE_TO_CODE:
        raise NotImplementedError(f"pixel type {arr.dtype} cannot be written as NIfTI")
    ncomp = 1
    if image.is_vector:
        ncomp = arr.shape[3]
        arr = np.ascontiguousarray(np.moveaxis(arr, -1, 0))  # planar in the file
    nz, ny, nx = arr.shape[-3:]
    spacing = np.asarray(image.GetSpacing(), dtype=np.float64)
    direction = np.asarray(image.GetDirection(), dtype=np.float64).reshape(3, 3)
    ras_rot = _LPS @ direction
    ras_org = _LPS @ np.asarray(image.GetOrigin(), dtype=np.float64)
    b, c, d, qfac = matrix_to_quaternion(ras_rot)
    srow = np.hstack([ras_rot * spacing[None, :], ras_org[:, None]])
    hdr = bytearray(348)
    struct.pack_into("<i", hdr, 0, 348)
    if ncomp == 1:
        struct.pack_into("<8h", hdr, 40, 3, nx, ny, nz, 1, 1, 1, 1)
    else:
        struct.pack_into("<8h", hdr, 40, 5, nx, ny, nz, 1, ncomp, 1, 1)
        struct.pack_into("<h", hdr, 68, 1007)  # intent_code: NIFTI_INTENT_VECTOR
    struct.pack_into("<h", hdr, 70, _DTYPE_TO_CODE[arr.dtype])
    struct.pack_into("<h", hdr, 72, arr.dtype.itemsize * 8)
    struct.pack_into("<8f", hdr, 76, qfac, spacing[0], spacing[1], spacing[2], 0.0, 1.0 if ncomp > 1 else 0.0, 0.0, 0.0)
    struct.pack_into("<fff", hdr, 108, 352.0, 1.0, 0.0)
    struct.pack_into("<B", hdr, 123, 2 | 8)  # xyzt_units: millimetres, seconds
    struct.pack_into("<hh", hdr, 252, 1, 1)   # qform_code, sform_code: NIFTI_XFORM_SCANNER_ANAT
    struct.pack_into("<6f", hdr, 256, b, c, d, ras_org[0], ras_org[1], ras_org[2])
    struct.pack_into("<12f", hdr, 280, *srow.reshape(12))
    hdr[344:348] = b"n+1\0"
    with _open(path, "wb") as f:
        f.write(bytes(hdr))
        f.write(b"\0\0\0\0")  # extension flag; data start at 352
        f.write(arr.astype(arr.dtype.newbyteorder("<"), copy=False).tobytes())
