"""
ctypes front-end of the CPU oracle (``oracle/itk_oracle.c``) -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

PARITY UNPINNED: see the header of ``itk_oracle.c``.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libitk_oracle.so")


def build(force=False):
    src = os.path.join(_HERE, "itk_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _LIB_PATH


class Geom(C.Structure):
    _fields_ = [("size", C.c_int32 * 3), ("spacing", C.c_double * 3), ("origin", C.c_double * 3), ("direction", C.c_double * 9)]


class TransformSpec(C.Structure):
    _fields_ = [("kind", C.c_int32), ("pad", C.c_int32), ("matrix", C.c_double * 9), ("offset", C.c_double * 3),
                ("dvf", C.c_void_p), ("dvf_geom", Geom)]


class DemonsParams(C.Structure):
    _fields_ = [("std_dev", C.c_double * 3), ("update_std_dev", C.c_double * 3),
                ("smooth_displacement_field", C.c_int32), ("smooth_update_field", C.c_int32),
                ("max_error", C.c_double), ("max_kernel_width", C.c_int32), ("number_of_iterations", C.c_int32),
                ("max_rms_error", C.c_double), ("max_update_step_length", C.c_double),
                ("intensity_difference_threshold", C.c_double), ("denominator_threshold", C.c_double)]


class DemonsStats(C.Structure):
    _fields_ = [("elapsed_iterations", C.c_int32), ("pad", C.c_int32), ("metric", C.c_double), ("rms_change", C.c_double)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_gaussian_operator.restype = C.c_int
        _lib.orc_gaussian_operator.argtypes = [C.c_double, C.c_double, C.c_int, C.c_void_p, C.c_int]
        _lib.orc_deriche_setup.restype = None
        _lib.orc_deriche_setup.argtypes = [C.c_double, C.c_double, C.c_void_p]
        _lib.orc_staple.argtypes = [C.c_void_p, C.c_int, C.c_size_t, C.c_double, C.c_uint, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.orc_rescale_threshold_f64.argtypes = [C.c_void_p, C.c_size_t, C.c_double]
        _lib.orc_combine_labels_f32.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_double, C.c_double, C.c_void_p]
        _lib.orc_resample_scalar.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_double]
        _lib.orc_resample_vec3.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_double]
        _lib.orc_discrete_gaussian_f32.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_int]
        _lib.orc_demons_execute.argtypes = [C.c_void_p] * 8
        _lib.orc_demons_force.argtypes = [C.c_void_p] * 10
        _lib.orc_pde_smooth_field.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_int]
        _lib.orc_recursive_gaussian_vec3.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.orc_transform_to_dvf.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        _lib.orc_set_num_threads.argtypes = [C.c_int]
        _lib.orc_set_semantic.argtypes = [C.c_char_p, C.c_int]
        _lib.orc_get_semantic.argtypes = [C.c_char_p]
        _lib.orc_semantic_name.restype = C.c_char_p
        _lib.orc_semantic_name.argtypes = [C.c_int]
        _lib.orc_semantic_meaning.restype = C.c_char_p
        _lib.orc_semantic_meaning.argtypes = [C.c_int]
        _lib.orc_bspline3_coefficients.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        _lib.orc_binary_fillhole.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        _lib.orc_largest_component.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.orc_signed_maurer.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        _lib.orc_label_contour.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        _lib.orc_label_contour.restype = None
        _lib.orc_binary_morph.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        _lib.orc_binary_morph.restype = None
    return _lib


def num_threads():
    return int(lib().orc_num_threads())


def set_num_threads(n):
    lib().orc_set_num_threads(int(n))


def semantics():
    """{name: (value, meaning)} of the named ITK-semantics switches (itk_oracle.c: g_semantics)."""
    l = lib()
    return {l.orc_semantic_name(i).decode(): (l.orc_get_semantic(l.orc_semantic_name(i)), l.orc_semantic_meaning(i).decode())
            for i in range(l.orc_num_semantics())}


def set_semantic(name, value):
    if lib().orc_set_semantic(name.encode(), int(value)) != 0:
        raise ValueError(f"unknown semantic switch {name!r}")


def get_semantic(name):
    v = lib().orc_get_semantic(name.encode())
    if v < 0:
        raise ValueError(f"unknown semantic switch {name!r}")
    return v


class semantic:
    """``with semantic(name, value): ...`` -- flips a switch for the block and restores it."""

    def __init__(self, name, value):
        self.name, self.value = name, value

    def __enter__(self):
        self.old = get_semantic(self.name)
        set_semantic(self.name, self.value)
        return self

    def __exit__(self, *exc):
        set_semantic(self.name, self.old)
        return False


def make_geom(size, spacing, origin, direction):
    g = Geom()
    for i in range(3):
        g.size[i] = int(size[i])
        g.spacing[i] = float(spacing[i])
        g.origin[i] = float(origin[i])
    for i in range(9):
        g.direction[i] = float(direction[i])
    return g


def geom_of(img):
    return make_geom(img.GetSize(), img.GetSpacing(), img.GetOrigin(), img.GetDirection())


_NP_TO_ORC = {np.dtype(np.int8): 0, np.dtype(np.uint8): 1, np.dtype(np.int16): 2, np.dtype(np.uint16): 3,
              np.dtype(np.int32): 4, np.dtype(np.uint32): 5, np.dtype(np.int64): 6, np.dtype(np.uint64): 7,
              np.dtype(np.float32): 8, np.dtype(np.float64): 9}


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def gaussian_operator(variance, max_error, max_width):
    cap = 2 * (max_width + 8) + 1
    buf = np.zeros(cap, dtype=np.float64)
    r = lib().orc_gaussian_operator(float(variance), float(max_error), int(max_width), _ptr(buf), cap)
    if r < 0:
        raise RuntimeError("gaussian operator overflow")
    return buf[: 2 * r + 1].copy()


def deriche_coefficients(sigma, spacing):
    buf = np.zeros(20, dtype=np.float64)
    lib().orc_deriche_setup(float(sigma), float(spacing), _ptr(buf))
    return buf


def _chain(transforms):
    """transforms: list in application order of ('affine', matrix9, offset3) / ('dvf', array[z,y,x,3] f64, Geom)."""
    n = len(transforms)
    arr = (TransformSpec * max(n, 1))()
    keep = []
    for i, t in enumerate(transforms):
        if t[0] == "affine":
            arr[i].kind = 0
            m = np.asarray(t[1], dtype=np.float64).reshape(9)
            o = np.asarray(t[2], dtype=np.float64).reshape(3)
            for k in range(9):
                arr[i].matrix[k] = m[k]
            for k in range(3):
                arr[i].offset[k] = o[k]
        else:
            arr[i].kind = 1
            d = np.ascontiguousarray(t[1], dtype=np.float64)
            keep.append(d)
            arr[i].dvf = d.ctypes.data
            arr[i].dvf_geom = t[2]
    return arr, n, keep


def resample_scalar(arr, gin, gout, transforms=(), interp=2, default_value=0.0):
    arr = np.ascontiguousarray(arr)
    out = np.empty((gout.size[2], gout.size[1], gout.size[0]), dtype=arr.dtype)
    ch, n, keep = _chain(list(transforms))
    rc = lib().orc_resample_scalar(_ptr(arr), _NP_TO_ORC[arr.dtype], C.byref(gin), _ptr(out), C.byref(gout), ch, n, int(interp), float(default_value))
    assert rc == 0
    return out


def resample_vec3(arr, gin, gout, transforms=(), default_value=0.0):
    arr = np.ascontiguousarray(arr, dtype=np.float64)
    out = np.empty((gout.size[2], gout.size[1], gout.size[0], 3), dtype=np.float64)
    ch, n, keep = _chain(list(transforms))
    rc = lib().orc_resample_vec3(_ptr(arr), C.byref(gin), _ptr(out), C.byref(gout), ch, n, float(default_value))
    assert rc == 0
    return out


def transform_to_dvf(gout, transforms):
    out = np.empty((gout.size[2], gout.size[1], gout.size[0], 3), dtype=np.float64)
    ch, n, keep = _chain(list(transforms))
    rc = lib().orc_transform_to_dvf(C.byref(gout), ch, n, _ptr(out))
    assert rc == 0
    return out


def discrete_gaussian_f32(arr, geom, variance, max_width=32, max_error=0.01, use_spacing=True):
    arr = np.ascontiguousarray(arr, dtype=np.float32)
    out = np.empty_like(arr)
    if np.isscalar(variance):
        variance = (variance,) * 3
    var = np.asarray(variance, dtype=np.float64)
    rc = lib().orc_discrete_gaussian_f32(_ptr(arr), _ptr(out), C.byref(geom), _ptr(var), int(max_width), float(max_error), int(bool(use_spacing)))
    if rc != 0:
        raise RuntimeError("discrete gaussian failed")
    return out


def demons_params(std_dev, iterations, update_std_dev=(1.0, 1.0, 1.0), smooth_displacement_field=True,
                  smooth_update_field=False, max_error=0.1, max_kernel_width=30, max_rms_error=0.02,
                  max_update_step_length=0.5, intensity_difference_threshold=0.001, denominator_threshold=1e-9):
    p = DemonsParams()
    for i in range(3):
        p.std_dev[i] = float(std_dev[i])
        p.update_std_dev[i] = float(update_std_dev[i])
    p.smooth_displacement_field = int(smooth_displacement_field)
    p.smooth_update_field = int(smooth_update_field)
    p.max_error = max_error
    p.max_kernel_width = max_kernel_width
    p.number_of_iterations = int(iterations)
    p.max_rms_error = max_rms_error
    p.max_update_step_length = max_update_step_length
    p.intensity_difference_threshold = intensity_difference_threshold
    p.denominator_threshold = denominator_threshold
    return p


def demons_execute(F, gF, M, gM, params, trace=False):
    F = np.ascontiguousarray(F, dtype=np.float32)
    M = np.ascontiguousarray(M, dtype=np.float32)
    D = np.empty(F.shape + (3,), dtype=np.float64)
    stats = DemonsStats()
    tr = np.zeros(max(params.number_of_iterations, 1), dtype=np.float64) if trace else None
    rc = lib().orc_demons_execute(_ptr(F), C.addressof(gF), _ptr(M), C.addressof(gM), C.addressof(params), _ptr(D), C.addressof(stats),
                                  _ptr(tr) if trace else None)
    assert rc == 0
    out = {"elapsed_iterations": stats.elapsed_iterations, "metric": stats.metric, "rms_change": stats.rms_change}
    if trace:
        out["metric_trace"] = tr[: stats.elapsed_iterations]
    return D, out


def demons_force(F, gF, M, gM, D, params):
    F = np.ascontiguousarray(F, dtype=np.float32)
    M = np.ascontiguousarray(M, dtype=np.float32)
    D = np.ascontiguousarray(D, dtype=np.float64)
    W = np.empty(F.shape, dtype=np.float32)
    U = np.empty(F.shape + (3,), dtype=np.float64)
    metric = C.c_double(0)
    rms = C.c_double(0)
    lib().orc_demons_force(_ptr(F), C.addressof(gF), _ptr(M), C.addressof(gM), _ptr(D), C.addressof(params), _ptr(W), _ptr(U),
                           C.addressof(metric), C.addressof(rms))
    return W, U, metric.value, rms.value


def pde_smooth_field(field, geom, sd, max_error=0.1, max_width=30):
    f = np.array(field, dtype=np.float64, order="C", copy=True)
    sdv = np.asarray(sd, dtype=np.float64)
    lib().orc_pde_smooth_field(_ptr(f), C.byref(geom), _ptr(sdv), float(max_error), int(max_width))
    return f


def recursive_gaussian_vec3(field, geom, sigma):
    f = np.array(field, dtype=np.float64, order="C", copy=True)
    s = np.asarray(sigma, dtype=np.float64)
    rc = lib().orc_recursive_gaussian_vec3(_ptr(f), C.byref(geom), _ptr(s))
    if rc != 0:
        raise RuntimeError("RecursiveGaussianImageFilter: the number of pixels along a direction is less than 4")
    return f


def combine_labels_f32(labels, weights, geom, smooth_variance=1.0, threshold=1e-4):
    n = len(labels)
    labels = [np.ascontiguousarray(l, dtype=np.uint8) for l in labels]
    weights = [np.ascontiguousarray(w, dtype=np.float32) for w in weights]
    lp = (C.c_void_p * n)(*[l.ctypes.data for l in labels])
    wp = (C.c_void_p * n)(*[w.ctypes.data for w in weights])
    out = np.empty(labels[0].shape, dtype=np.float32)
    lib().orc_combine_labels_f32(lp, wp, n, C.byref(geom), float(smooth_variance), float(threshold or 0.0), _ptr(out))
    return out


def staple(decisions, confidence_weight=1.0, max_iter=0xFFFFFFFF):
    n = len(decisions)
    decisions = [np.ascontiguousarray(d, dtype=np.uint8) for d in decisions]
    dp = (C.c_void_p * n)(*[d.ctypes.data for d in decisions])
    W = np.empty(decisions[0].shape, dtype=np.float64)
    p = np.zeros(n)
    q = np.zeros(n)
    it = lib().orc_staple(dp, n, decisions[0].size, float(confidence_weight), C.c_uint(max_iter), _ptr(W), _ptr(p), _ptr(q))
    return W, p, q, it


def rescale_threshold_f64(img, threshold):
    out = np.array(img, dtype=np.float64, order="C", copy=True)
    lib().orc_rescale_threshold_f64(_ptr(out), out.size, float(threshold or 0.0))
    return out


def binary_fillhole(mask):
    """sitk.BinaryFillhole (face connectivity, foreground 1) on a [z, y, x] uint8 array."""
    m = np.ascontiguousarray(mask, dtype=np.uint8)
    out = np.empty_like(m)
    nz, ny, nx = m.shape
    if lib().orc_binary_fillhole(_ptr(m), nx, ny, nz, _ptr(out)) != 0:
        raise MemoryError
    return out


def largest_component(mask, want_labels=False):
    """ConnectedComponent -> largest object (first in raster order on ties) as uint8; also returns the object count,
    the size of the largest and (optionally) the ConnectedComponent label image."""
    m = np.ascontiguousarray(mask, dtype=np.uint8)
    out = np.empty_like(m)
    nz, ny, nx = m.shape
    labels = np.empty(m.shape, dtype=np.int32) if want_labels else None
    nvox = C.c_int64(0)
    ncomp = lib().orc_largest_component(_ptr(m), nx, ny, nz, _ptr(out), _ptr(labels) if want_labels else None, C.byref(nvox))
    if ncomp < 0:
        raise MemoryError
    return (out, ncomp, nvox.value, labels) if want_labels else (out, ncomp, nvox.value)


def bspline3_coefficients(arr, geom):
    """itk::BSplineDecompositionImageFilter (order 3) coefficients of a scalar [z, y, x] array, as float64."""
    a = np.ascontiguousarray(arr)
    out = np.empty(a.shape, dtype=np.float64)
    if lib().orc_bspline3_coefficients(_ptr(a), _NP_TO_ORC[a.dtype], C.byref(geom), _ptr(out)) != 0:
        raise MemoryError
    return out


def signed_maurer_distance_map(mask, spacing=(1.0, 1.0, 1.0), inside_is_positive=False, squared_distance=False, use_image_spacing=True):
    """sitk.SignedMaurerDistanceMap on a [z, y, x] mask (background 0) -> float32; ``spacing`` is (x, y, z)."""
    m = np.ascontiguousarray(mask != 0, dtype=np.uint8)
    out = np.empty(m.shape, dtype=np.float32)
    nz, ny, nx = m.shape
    sp = np.ascontiguousarray(spacing, dtype=np.float64)
    if lib().orc_signed_maurer(_ptr(m), nx, ny, nz, _ptr(sp), int(bool(inside_is_positive)), int(bool(squared_distance)),
                               int(bool(use_image_spacing)), _ptr(out)) != 0:
        raise MemoryError
    return out


def label_contour(mask, fully_connected=False):
    """sitk.LabelContour on a [z, y, x] uint8 label array (background 0)."""
    m = np.ascontiguousarray(mask, dtype=np.uint8)
    out = np.empty_like(m)
    nz, ny, nx = m.shape
    lib().orc_label_contour(_ptr(m), nx, ny, nz, int(bool(fully_connected)), _ptr(out))
    return out


def binary_morph(mask, offsets, dilate, boundary_to_foreground=None):
    """sitk.BinaryDilate / sitk.BinaryErode (foreground 1) with the structuring element given as (dx, dy, dz) offsets."""
    m = np.ascontiguousarray(mask, dtype=np.uint8)
    out = np.empty_like(m)
    nz, ny, nx = m.shape
    offs = np.ascontiguousarray(offsets, dtype=np.int32).reshape(-1, 3)
    if boundary_to_foreground is None:
        boundary_to_foreground = not dilate  # SimpleITK defaults
    lib().orc_binary_morph(_ptr(m), nx, ny, nz, _ptr(offs), int(offs.shape[0]), int(bool(dilate)), int(bool(boundary_to_foreground)), _ptr(out))
    return out
