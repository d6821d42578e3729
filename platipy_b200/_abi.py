"""
ctypes binding of ``libb200reg.so`` (C ABI declared in ``include/b200reg.h``).

The library is loaded from the package directory (built in-tree by ``__graft_entry__.build()`` /
``make -C platipy_b200/csrc``).  There is no fallback: a missing library raises ``ImportError`` and a
failing call raises the exception class the reference would raise for the same condition
(SimpleITK ``RuntimeError``, platipy ``ValueError`` / ``AttributeError``).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PLATIPY_B200_LIB") or os.environ.get("B200REG_LIB") or os.path.join(_HERE, "libb200reg.so")  # A/B testing of builds

MAX_TRANSFORMS = 4
MAX_LEVELS = 8
MAX_BATCH = 64

OK, ERR_CUDA, ERR_ARG, ERR_UNSUPPORTED, ERR_RUNTIME = 0, 1, 2, 3, 4
TFM_AFFINE, TFM_DVF = 0, 1
OP_OR, OP_AND, OP_ADD, OP_XOR = 0, 1, 2, 3


class Geom(C.Structure):
    _fields_ = [("size", C.c_int32 * 3), ("spacing", C.c_double * 3), ("origin", C.c_double * 3), ("direction", C.c_double * 9)]


class Transform(C.Structure):
    _fields_ = [("kind", C.c_int32), ("pad", C.c_int32), ("matrix", C.c_double * 9), ("offset", C.c_double * 3),
                ("d_dvf", C.c_void_p), ("dvf_geom", Geom)]


class DemonsParams(C.Structure):
    _fields_ = [("std_dev", C.c_double * 3), ("update_std_dev", C.c_double * 3),
                ("smooth_displacement_field", C.c_int32), ("smooth_update_field", C.c_int32),
                ("max_error", C.c_double), ("max_kernel_width", C.c_int32), ("number_of_iterations", C.c_int32),
                ("max_rms_error", C.c_double), ("max_update_step_length", C.c_double),
                ("intensity_difference_threshold", C.c_double), ("denominator_threshold", C.c_double),
                ("field_precision", C.c_int32), ("reserved", C.c_int32)]


class DemonsStats(C.Structure):
    _fields_ = [("elapsed_iterations", C.c_int32), ("voxels_lo", C.c_int32), ("metric", C.c_double),
                ("rms_change", C.c_double), ("gpu_ms", C.c_double)]


class MultiresConfig(C.Structure):
    _fields_ = [("n_levels", C.c_int32), ("isotropic_resample", C.c_int32),
                ("resolution_staging", C.c_double * MAX_LEVELS), ("smoothing_sigmas", C.c_double * MAX_LEVELS),
                ("iteration_staging", C.c_int32 * MAX_LEVELS), ("interp_order", C.c_int32), ("demons", DemonsParams)]


# every symbol include/b200reg.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SIGNATURES = {
    "b200reg_abi_version": (C.c_int, []),
    "b200reg_last_error": (C.c_char_p, []),
    "b200reg_create": (C.c_int, [C.c_int, _P, C.POINTER(_P)]),
    "b200reg_destroy": (C.c_int, [_P]),
    "b200reg_set_stream": (C.c_int, [_P, _P]),
    "b200reg_synchronize": (C.c_int, [_P]),
    "b200reg_launch_count": (C.c_int64, [_P]),
    "b200reg_set_semantic": (C.c_int, [C.c_char_p, C.c_int]),
    "b200reg_get_semantic": (C.c_int, [C.c_char_p]),
    "b200reg_malloc": (C.c_int, [_P, C.c_size_t, C.POINTER(_P)]),
    "b200reg_free": (C.c_int, [_P, _P]),
    "b200reg_malloc_host": (C.c_int, [C.c_size_t, C.POINTER(_P)]),
    "b200reg_free_host": (C.c_int, [_P]),
    "b200reg_memcpy_h2d": (C.c_int, [_P, _P, _P, C.c_size_t]),
    "b200reg_memcpy_d2h": (C.c_int, [_P, _P, _P, C.c_size_t]),
    "b200reg_memset": (C.c_int, [_P, _P, C.c_int, C.c_size_t]),
    "b200reg_aos_to_soa": (C.c_int, [_P, _P, _P, C.c_size_t]),
    "b200reg_soa_to_aos": (C.c_int, [_P, _P, _P, C.c_size_t]),
    "b200reg_cast": (C.c_int, [_P, _P, C.c_int, _P, C.c_int, C.c_size_t]),
    "b200reg_minmax": (C.c_int, [_P, _P, C.c_int, C.c_size_t, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "b200reg_discrete_gaussian_f32": (C.c_int, [_P, _P, _P, C.POINTER(Geom), C.POINTER(C.c_double), C.c_int, C.c_double, C.c_int]),
    "b200reg_identity_resample_is_exact": (C.c_int, [C.POINTER(Geom), C.POINTER(Geom), C.c_int]),
    "b200reg_smooth_and_resample_f32": (C.c_int, [_P, _P, C.POINTER(Geom), C.POINTER(C.c_double), C.c_int, C.POINTER(Geom), C.c_int, _P, C.c_int]),
    "b200reg_gaussian_operator": (C.c_int, [C.c_double, C.c_double, C.c_int, C.POINTER(C.c_double), C.c_int]),
    "b200reg_resample": (C.c_int, [_P, _P, C.c_int, C.POINTER(Geom), _P, C.POINTER(Geom), C.POINTER(Transform), C.c_int, C.c_int, C.c_double]),
    "b200reg_resample_batch": (C.c_int, [_P, C.c_int, C.POINTER(_P), C.POINTER(C.c_int), C.POINTER(Geom), C.POINTER(_P), C.POINTER(Geom),
                                         C.POINTER(Transform), C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_double)]),
    "b200reg_resample_vec3": (C.c_int, [_P, _P, C.POINTER(Geom), _P, C.POINTER(Geom), C.POINTER(Transform), C.c_int, C.c_double]),
    "b200reg_transform_to_dvf": (C.c_int, [_P, C.POINTER(Geom), C.POINTER(Transform), C.c_int, _P]),
    "b200reg_compose_dvf": (C.c_int, [_P, _P, _P, C.POINTER(Geom), _P]),
    "b200reg_demons_execute": (C.c_int, [_P, _P, C.POINTER(Geom), _P, C.POINTER(Geom), C.POINTER(DemonsParams), _P, C.POINTER(DemonsStats)]),
    "b200reg_demons_trace": (C.c_int, [_P, C.c_int, C.POINTER(C.c_double), C.c_int]),
    "b200reg_demons_force": (C.c_int, [_P, _P, C.POINTER(Geom), _P, C.POINTER(Geom), _P, C.POINTER(DemonsParams), _P, _P,
                                       C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "b200reg_pde_smooth_field": (C.c_int, [_P, _P, C.POINTER(Geom), C.POINTER(C.c_double), C.c_double, C.c_int]),
    "b200reg_recursive_gaussian_vec3": (C.c_int, [_P, _P, C.POINTER(Geom), C.POINTER(C.c_double)]),
    "b200reg_multiscale_demons": (C.c_int, [_P, _P, C.POINTER(Geom), _P, C.POINTER(Geom), C.POINTER(MultiresConfig), _P, C.POINTER(Geom), _P,
                                            C.POINTER(DemonsStats)]),
    "b200reg_pyramid_geom": (C.c_int, [C.POINTER(Geom), C.c_int, C.c_double, C.POINTER(Geom)]),
    "b200reg_weight_map": (C.c_int, [_P, _P, _P, C.POINTER(Geom), C.c_int, C.c_double, C.c_double, C.c_double, _P]),
    "b200reg_weight_map_block": (C.c_int, [_P, _P, _P, C.POINTER(Geom), C.POINTER(C.c_int32), C.c_double, C.c_double, _P]),
    "b200reg_normalise_by_max": (C.c_int, [_P, _P, _P, C.c_size_t]),
    "b200reg_vote_accumulate": (C.c_int, [_P, _P, _P, _P, _P, C.c_size_t, C.c_int]),
    "b200reg_vote_finalize": (C.c_int, [_P, _P, _P, C.POINTER(Geom), C.c_double, C.c_double, _P]),
    "b200reg_binary_threshold": (C.c_int, [_P, _P, C.c_int, C.c_size_t, C.c_double, C.c_double, _P]),
    "b200reg_pack_decision": (C.c_int, [_P, _P, C.c_int, _P, C.c_size_t, C.c_int]),
    "b200reg_unpack_decision": (C.c_int, [_P, _P, C.c_int, _P, C.c_size_t]),
    "b200reg_pack_label": (C.c_int, [_P, _P, C.c_int, _P, C.c_int, C.c_size_t, C.c_int]),
    "b200reg_staple_packed": (C.c_int, [_P, _P, C.c_int, C.c_uint32, C.c_size_t, C.c_double, C.c_uint32, C.c_double, C.c_int, _P,
                                        C.POINTER(C.c_double), C.POINTER(C.c_int32)]),
    "b200reg_count_accumulate": (C.c_int, [_P, _P, _P, C.c_size_t, C.c_int, _P]),
    "b200reg_vote_finalize_counts": (C.c_int, [_P, _P, C.c_int, C.POINTER(Geom), C.c_double, C.c_double, _P]),
    "b200reg_binary_fillhole": (C.c_int, [_P, _P, C.POINTER(C.c_int32), C.c_int, _P]),
    "b200reg_largest_component": (C.c_int, [_P, _P, C.POINTER(C.c_int32), C.c_int, _P, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "b200reg_process_probability": (C.c_int, [_P, _P, C.c_int, C.POINTER(C.c_int32), C.c_double, _P, C.POINTER(C.c_int64)]),
    "b200reg_linreg_meansq": (C.c_int, [_P, _P, C.POINTER(Geom), _P, C.POINTER(Geom), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                        C.POINTER(C.c_double), C.POINTER(C.c_double), _P, _P, C.c_int, C.POINTER(C.c_double)]),
    "b200reg_linreg_correlation": (C.c_int, [_P, _P, C.POINTER(Geom), _P, C.POINTER(Geom), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                             C.POINTER(C.c_double), C.POINTER(C.c_double), _P, _P, C.c_int, C.POINTER(C.c_double)]),
    "b200reg_linreg_mattes_histogram": (C.c_int, [_P, _P, C.POINTER(Geom), _P, C.POINTER(Geom), C.POINTER(C.c_double), C.POINTER(C.c_double), _P, _P,
                                                  C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                                  C.POINTER(C.c_double)]),
    "b200reg_linreg_mattes_derivative": (C.c_int, [_P, _P, C.POINTER(Geom), _P, C.POINTER(Geom), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                                   C.POINTER(C.c_double), C.POINTER(C.c_double), _P, _P, C.c_int, C.c_int, C.POINTER(C.c_double),
                                                   C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "b200reg_image_moments": (C.c_int, [_P, _P, C.POINTER(Geom), C.POINTER(C.c_double)]),
    "b200reg_bounding_box": (C.c_int, [_P, _P, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "b200reg_region_copy": (C.c_int, [_P, _P, C.POINTER(C.c_int32), C.POINTER(C.c_int32), _P, C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                      C.POINTER(C.c_int32), C.c_int]),
    "b200reg_resolve_overlap": (C.c_int, [_P, C.POINTER(_P), C.POINTER(_P), C.c_int, C.c_size_t]),
    "b200reg_binary_closing": (C.c_int, [_P, _P, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int, _P]),
    "b200reg_staple": (C.c_int, [_P, C.POINTER(_P), C.c_int, C.c_size_t, C.c_double, C.c_uint32, C.c_double, C.c_int, _P,
                                 C.POINTER(C.c_double), C.POINTER(C.c_int32)]),
    "b200reg_signed_maurer_distance_map": (C.c_int, [_P, _P, C.POINTER(Geom), C.c_int, C.c_int, C.c_int, _P]),
    "b200reg_label_contour": (C.c_int, [_P, _P, C.POINTER(C.c_int32), C.c_int, _P]),
    "b200reg_label_contour_slicewise": (C.c_int, [_P, _P, C.POINTER(C.c_int32), _P]),
    "b200reg_binary_dilate": (C.c_int, [_P, _P, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int, C.c_int, _P]),
    "b200reg_binary_erode": (C.c_int, [_P, _P, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int, C.c_int, _P]),
    "b200reg_u8_binary_op": (C.c_int, [_P, _P, _P, C.c_int, _P, C.c_size_t]),
    "b200reg_mask_image": (C.c_int, [_P, _P, C.c_int, _P, C.c_size_t, C.c_int, C.c_double, _P]),
    "b200reg_divide_scalar": (C.c_int, [_P, _P, C.c_int, C.c_size_t, C.c_double, _P]),
    "b200reg_constant_field": (C.c_int, [_P, _P, C.c_size_t, C.POINTER(C.c_double), _P]),
    "b200reg_patch_correlation": (C.c_int, [_P, _P, _P, C.POINTER(C.c_int32), C.POINTER(C.c_int32), _P]),
    "b200reg_scale_shift": (C.c_int, [_P, _P, C.c_int, C.c_size_t, C.c_int, C.c_double, C.c_double, _P]),
    "b200reg_radial_bend_field": (C.c_int, [_P, _P, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_double), C.c_double, C.c_int,
                                            C.c_int, _P]),
}

_lib = None


def load():
    """Load libb200reg.so and bind every declared symbol; raises ImportError when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C platipy_b200/csrc`).  platipy_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.b200reg_abi_version() != 1:
        raise ImportError("libb200reg.so ABI version mismatch")
    _lib = lib
    return lib


SEMANTIC_SWITCHES = ("discrete_gaussian_axis_order", "recursive_gaussian_axis_order", "resample_linear_scanline", "dvf_transform_interpolation",
                     "vector_resample_interpolation", "binary_threshold_in_pixel_type")


def set_semantic(name, value):
    """Flip one of the named ITK-semantics switches of the library (process-wide; see include/b200reg.h)."""
    check(load().b200reg_set_semantic(name.encode(), int(value)))


def get_semantic(name):
    v = load().b200reg_get_semantic(name.encode())
    if v < 0:
        raise ValueError(f"unknown semantic switch {name!r}")
    return v


class B200RegNotImplemented(NotImplementedError):
    pass


def check(status):
    """Map a b200reg_status to the exception class the reference raises for the same condition."""
    if status == OK:
        return
    msg = load().b200reg_last_error().decode("utf-8", "replace")
    if status == ERR_ARG:
        raise ValueError(msg)
    if status == ERR_UNSUPPORTED:
        raise B200RegNotImplemented(msg)
    raise RuntimeError(msg)  # CUDA and ITK-style runtime failures (SimpleITK raises RuntimeError)


def make_geom(size, spacing, origin, direction):
    g = Geom()
    for i in range(3):
        g.size[i] = int(size[i])
        g.spacing[i] = float(spacing[i])
        g.origin[i] = float(origin[i])
    for i in range(9):
        g.direction[i] = float(direction[i])
    return g


def geom_of(image):
    return make_geom(image.GetSize(), image.GetSpacing(), image.GetOrigin(), image.GetDirection())


def geom_tuple(g):
    return (tuple(g.size), tuple(g.spacing), tuple(g.origin), tuple(g.direction))
