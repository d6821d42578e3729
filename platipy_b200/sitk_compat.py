"""
Image / transform containers with the SimpleITK surface the hot path touches.

The reference functions take and return ``SimpleITK.Image`` / ``SimpleITK.Transform`` objects
(reference platipy/imaging/registration/deformable.py:190-306, utils.py:148-267).  SimpleITK is not
installable in this environment, so this module provides value-semantic stand-ins exposing the same
method names (``GetSize``, ``GetSpacing``, ``GetOrigin``, ``GetDirection``, ``GetPixelID``,
``CopyInformation`` ...) and the module-level helpers (``GetArrayFromImage``, ``GetImageFromArray``,
``Cast``) the callers use.  When a real SimpleITK *is* importable, ``to_native``/``from_native`` convert
at the boundary so the public functions accept and return real ``SimpleITK.Image`` objects.

numpy layout is ``[z, y, x]`` (+ trailing component axis for vector images), ``GetSize()`` is ``(x, y, z)``.
"""
from __future__ import annotations

import numpy as np

try:  # pragma: no cover - SimpleITK is absent in the build container
    import SimpleITK as _sitk

    HAVE_SITK = True
except Exception:  # noqa: BLE001
    _sitk = None
    HAVE_SITK = False

# SimpleITK pixel IDs (Linux builds; reference SURVEY Appendix A.1)
sitkUnknown = -1
sitkInt8, sitkUInt8, sitkInt16, sitkUInt16, sitkInt32, sitkUInt32, sitkInt64, sitkUInt64 = range(8)
sitkFloat32, sitkFloat64 = 8, 9
sitkVectorFloat32, sitkVectorFloat64 = 20, 21

# interpolator enum (reference deformable.py:221-224)
sitkNearestNeighbor, sitkLinear, sitkBSpline = 1, 2, 3
# event enum of AddCommand (reference deformable.py:261: sitk.sitkIterationEvent)
sitkAnyEvent, sitkAbortEvent, sitkDeleteEvent, sitkEndEvent, sitkIterationEvent, sitkProgressEvent, sitkStartEvent = range(7)

_ID_TO_DTYPE = {
    sitkInt8: np.int8, sitkUInt8: np.uint8, sitkInt16: np.int16, sitkUInt16: np.uint16,
    sitkInt32: np.int32, sitkUInt32: np.uint32, sitkInt64: np.int64, sitkUInt64: np.uint64,
    sitkFloat32: np.float32, sitkFloat64: np.float64,
    sitkVectorFloat32: np.float32, sitkVectorFloat64: np.float64,
}
_DTYPE_TO_ID = {np.dtype(v): k for k, v in _ID_TO_DTYPE.items() if k < 10}


def pixel_id_to_dtype(pixel_id):
    return np.dtype(_ID_TO_DTYPE[int(pixel_id)])


def dtype_to_pixel_id(dtype, is_vector=False):
    dtype = np.dtype(dtype)
    if is_vector:
        return {np.dtype(np.float32): sitkVectorFloat32, np.dtype(np.float64): sitkVectorFloat64}[dtype]
    if dtype == np.bool_:
        return sitkUInt8
    return _DTYPE_TO_ID[dtype]


class Image:
    """Stand-in for ``SimpleITK.Image`` (3-D scalar or 3-component vector)."""

    __slots__ = ("_arr", "_spacing", "_origin", "_direction", "_is_vector")

    def __init__(self, array, spacing=(1.0, 1.0, 1.0), origin=(0.0, 0.0, 0.0),
                 direction=(1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0), is_vector=False):
        arr = np.asarray(array)
        if arr.dtype == np.bool_:
            arr = arr.astype(np.uint8)
        if is_vector:
            if arr.ndim != 4:
                raise ValueError("vector image needs a [z, y, x, c] array")
        elif arr.ndim != 3:
            raise ValueError("scalar image needs a [z, y, x] array")
        self._arr = np.ascontiguousarray(arr)
        self._spacing = tuple(float(s) for s in spacing)
        self._origin = tuple(float(s) for s in origin)
        self._direction = tuple(float(s) for s in direction)
        self._is_vector = bool(is_vector)

    # -- geometry ---------------------------------------------------------------------------------
    def GetSize(self):
        z, y, x = self._arr.shape[:3]
        return (int(x), int(y), int(z))

    def GetWidth(self):
        return self.GetSize()[0]

    def GetHeight(self):
        return self.GetSize()[1]

    def GetDepth(self):
        return self.GetSize()[2]

    def GetDimension(self):
        return 3

    def GetSpacing(self):
        return self._spacing

    def GetOrigin(self):
        return self._origin

    def GetDirection(self):
        return self._direction

    def SetSpacing(self, spacing):
        self._spacing = tuple(float(s) for s in spacing)

    def SetOrigin(self, origin):
        self._origin = tuple(float(s) for s in origin)

    def SetDirection(self, direction):
        self._direction = tuple(float(s) for s in direction)

    def CopyInformation(self, other):
        if tuple(other.GetSize()) != self.GetSize():
            raise RuntimeError("CopyInformation: source image size does not match this image's size")
        self._spacing = tuple(other.GetSpacing())
        self._origin = tuple(other.GetOrigin())
        self._direction = tuple(other.GetDirection())

    # -- pixels -----------------------------------------------------------------------------------
    def GetPixelID(self):
        return dtype_to_pixel_id(self._arr.dtype, self._is_vector)

    def GetPixelIDValue(self):
        return self.GetPixelID()

    def GetNumberOfComponentsPerPixel(self):
        return int(self._arr.shape[3]) if self._is_vector else 1

    def GetNumberOfPixels(self):
        z, y, x = self._arr.shape[:3]
        return int(x) * int(y) * int(z)

    def __len__(self):
        return self.GetNumberOfPixels()

    def __bool__(self):
        return True

    @property
    def array(self):
        return self._arr

    @property
    def is_vector(self):
        return self._is_vector

    # image-with-constant operators of SimpleITK's Image, for user callbacks such as compute_weight_map's
    # ``correlation_function`` (fusion.py:138-146).  Host numpy on the container; the product's own paths never use them.
    def _with(self, arr):
        return Image(np.asarray(arr, dtype=self._arr.dtype), self._spacing, self._origin, self._direction, self._is_vector)

    def __add__(self, c):
        return self._with(self._arr + self._arr.dtype.type(c)) if np.isscalar(c) else NotImplemented

    __radd__ = __add__

    def __sub__(self, c):
        return self._with(self._arr - self._arr.dtype.type(c)) if np.isscalar(c) else NotImplemented

    def __rsub__(self, c):
        return self._with(self._arr.dtype.type(c) - self._arr) if np.isscalar(c) else NotImplemented

    def __mul__(self, c):
        return self._with(self._arr * self._arr.dtype.type(c)) if np.isscalar(c) else NotImplemented

    __rmul__ = __mul__

    def __truediv__(self, c):
        return self._with(self._arr / self._arr.dtype.type(c)) if np.isscalar(c) else NotImplemented

    def __neg__(self):
        return self._with(-self._arr)

    def __abs__(self):
        return self._with(np.abs(self._arr))

    def same_space(self, other, tol=1e-6):
        return (
            self.GetSize() == tuple(other.GetSize())
            and np.allclose(self._spacing, other.GetSpacing(), rtol=0, atol=tol)
            and np.allclose(self._origin, other.GetOrigin(), rtol=0, atol=tol)
            and np.allclose(self._direction, other.GetDirection(), rtol=0, atol=tol)
        )


def GetArrayFromImage(image):
    image = to_native(image)
    return image.array.copy()


def GetArrayViewFromImage(image):
    image = to_native(image)
    view = image.array.view()
    view.flags.writeable = False
    return view


def GetImageFromArray(arr, isVector=None):
    arr = np.asarray(arr)
    if isVector is None:
        isVector = arr.ndim == 4
    return Image(arr.copy(), is_vector=bool(isVector))


def Cast(image, pixel_id):
    """``sitk.Cast``: C ``static_cast`` semantics (float -> int truncates toward zero)."""
    image = to_native(image)
    dtype = pixel_id_to_dtype(pixel_id)
    src = image.array
    if src.dtype == dtype:
        out = src.copy()
    elif np.issubdtype(dtype, np.integer) and np.issubdtype(src.dtype, np.floating):
        out = np.trunc(src).astype(dtype)
    else:
        out = src.astype(dtype)
    return Image(out, image.GetSpacing(), image.GetOrigin(), image.GetDirection(), image.is_vector)


def ReadImage(path, outputPixelType=sitkUnknown):
    """``sitk.ReadImage``: real SimpleITK when it is importable, otherwise the NIfTI-1 reader of ``nifti_io``."""
    if HAVE_SITK:  # pragma: no cover
        return _sitk.ReadImage(str(path)) if outputPixelType == sitkUnknown else _sitk.ReadImage(str(path), outputPixelType)
    from . import nifti_io

    img = nifti_io.read_image(path)
    return img if outputPixelType == sitkUnknown else Cast(img, outputPixelType)


def WriteImage(image, path):
    """``sitk.WriteImage`` (NIfTI-1 without SimpleITK)."""
    if HAVE_SITK and isinstance(image, _sitk.Image):  # pragma: no cover
        return _sitk.WriteImage(image, str(path))
    from . import nifti_io

    nifti_io.write_image(image, path)


# -- transforms -----------------------------------------------------------------------------------
class Transform:
    """``sitk.Transform()``: identity."""

    def IsLinear(self):
        return True

    def flatten(self):
        """Transforms in application order (first applied first)."""
        return []


class AffineTransform(Transform):
    """``p -> M (p - c) + c + t`` (itk::MatrixOffsetTransformBase); covers the rigid/similarity family."""

    def __init__(self, matrix=(1, 0, 0, 0, 1, 0, 0, 0, 1), translation=(0, 0, 0), center=(0, 0, 0)):
        self.matrix = np.asarray(matrix, dtype=np.float64).reshape(3, 3)
        self.translation = np.asarray(translation, dtype=np.float64).reshape(3)
        self.center = np.asarray(center, dtype=np.float64).reshape(3)

    @property
    def offset(self):
        # MatrixOffsetTransformBase::ComputeOffset: offset_i = t_i + c_i - sum_j M_ij c_j
        off = np.empty(3)
        for i in range(3):
            off[i] = self.translation[i] + self.center[i]
            for j in range(3):
                off[i] -= self.matrix[i, j] * self.center[j]
        return off

    def flatten(self):
        return [self]

    def TransformPoint(self, p):
        p = np.asarray(p, dtype=np.float64)
        return tuple(self.matrix @ p + self.offset)


class DisplacementFieldTransform(Transform):
    """``sitk.DisplacementFieldTransform(image)``: ``p -> p + D(p)``, identity outside the field buffer.

    Like SimpleITK the constructor takes ownership of the field (reference deformable.py:139,296 always
    pass a fresh ``sitk.Cast`` copy).  A device-resident copy of the field is cached on the object by
    ``platipy_b200.registration`` so repeated ``apply_transform`` calls do not re-upload it
    (reference multiatlas/run.py:331-345 calls apply_transform once per structure).
    """

    def __init__(self, image):
        image = to_native(image)
        if not image.is_vector or image.array.shape[3] != 3:
            raise RuntimeError("DisplacementFieldTransform needs a 3-component vector image")
        if image.array.dtype != np.float64:
            raise RuntimeError("DisplacementFieldTransform needs a sitkVectorFloat64 image")
        self._field = image
        self._device_cache = None

    def IsLinear(self):
        return False

    def GetDisplacementField(self):
        return self._field

    def flatten(self):
        return [self]


class CompositeTransform(Transform):
    """``sitk.CompositeTransform([T1, T2])``: ``p -> T1(T2(p))`` (last added is applied first)."""

    def __init__(self, transforms=()):
        if isinstance(transforms, Transform):
            transforms = [transforms]
        self.transforms = list(transforms)

    def AddTransform(self, t):
        self.transforms.append(t)

    def IsLinear(self):
        return all(t.IsLinear() for t in self.transforms)

    def flatten(self):
        out = []
        for t in reversed(self.transforms):
            out.extend(t.flatten())
        return out


def is_native_sitk(obj):
    """True for a real ``SimpleITK.Image`` (only when SimpleITK is importable)."""
    return bool(HAVE_SITK and isinstance(obj, _sitk.Image))


# -- boundary conversion -----------------------------------------------------------------------------
def to_native(image):
    """Accept a stand-in ``Image`` or a real ``SimpleITK.Image``; return a stand-in ``Image``."""
    if isinstance(image, Image):
        return image
    if HAVE_SITK and isinstance(image, _sitk.Image):  # pragma: no cover
        arr = _sitk.GetArrayFromImage(image)
        is_vec = image.GetNumberOfComponentsPerPixel() > 1
        return Image(arr, image.GetSpacing(), image.GetOrigin(), image.GetDirection(), is_vec)
    raise TypeError(f"expected an Image, got {type(image)!r}")


def from_native(image, like):
    """Return ``image`` as the same kind of object as ``like`` (real SimpleITK image if ``like`` is one)."""
    if HAVE_SITK and isinstance(like, _sitk.Image):  # pragma: no cover
        out = _sitk.GetImageFromArray(image.array, isVector=image.is_vector)
        out.SetSpacing(image.GetSpacing())
        out.SetOrigin(image.GetOrigin())
        out.SetDirection(image.GetDirection())
        return out
    return image
