"""
CPU restatement of the reference's synthetic-deformation generators and distance-map helpers -- TEST INFRASTRUCTURE, NOT
PRODUCT CODE.  PARITY UNPINNED (see ``itk_oracle.c``): every SimpleITK call is replaced by the oracle's restatement of the
ITK filter behind it.  Paths relative to /root/reference:

  convert_mask_to_distance_map / convert_mask_to_reg_structure    platipy/imaging/registration/utils.py:270-344
  generate_field_shift / asymmetric_contract / asymmetric_extend / expand / radial_bend   platipy/imaging/generation/dvf.py:29-415
  evaluate_distance_to_reference                                   platipy/imaging/label/projection.py:67-92

Images are ``platipy_b200.sitk_compat.Image`` containers (data holders only).
"""
from __future__ import annotations

import numpy as np

from platipy_b200 import sitk_compat as sk
from platipy_b200.sitk_compat import Image

from . import itk_oracle as orc
from . import platipy_ref as ref


def _like(arr, like, is_vector=False):
    return Image(arr, like.GetSpacing(), like.GetOrigin(), like.GetDirection(), is_vector)


def ball(radius):
    """itk::FlatStructuringElement::Ball(radius) as (dx, dy, dz) offsets: voxel centres inside the ellipsoid with semi-axes
    r + 1/2 [ITK-recall, the symmetric form; see DESIGN.md section 5]."""
    rx, ry, rz = (int(v) for v in radius)
    offs = []
    for dz in range(-rz, rz + 1):
        for dy in range(-ry, ry + 1):
            for dx in range(-rx, rx + 1):
                if (dx / (rx + 0.5)) ** 2 + (dy / (ry + 0.5)) ** 2 + (dz / (rz + 0.5)) ** 2 <= 1.0:
                    offs.append((dx, dy, dz))
    return np.array(offs, dtype=np.int32)


def binary_dilate(mask, radius):
    return _like(orc.binary_morph(mask.array, ball(radius), True), mask)


def binary_erode(mask, radius):
    return _like(orc.binary_morph(mask.array, ball(radius), False), mask)


def _median_threshold(mask):
    # utils.py:282-287
    arr = mask.array
    vals = np.unique(arr[arr > 0])
    if len(vals) > 2:
        cutoff = np.median(vals)
        mask = _like(((arr >= cutoff) & (arr <= np.max(vals).astype(float))).astype(np.uint8), mask)
    return mask


def convert_mask_to_distance_map(mask, squared_distance=False, normalise=False):
    mask = _median_threshold(mask)
    raw = orc.signed_maurer_distance_map(mask.array, mask.GetSpacing(), inside_is_positive=True, squared_distance=squared_distance,
                                         use_image_spacing=True)
    if normalise:
        raw = raw / np.float32(raw.max())
    return _like(raw, mask)


def convert_mask_to_reg_structure(mask, expansion=(0, 0, 0), scale=lambda x: x):
    mask = _median_threshold(mask)
    if not hasattr(expansion, "__iter__"):
        expansion = [int(expansion / i) for i in mask.GetSpacing()]
    if any(expansion):
        mask = binary_dilate(mask, expansion)
    distance_map = convert_mask_to_distance_map(mask, squared_distance=False).array.astype(np.float64)
    distance_map = np.where(mask.array != 0, distance_map, 0.0)  # sitk.Mask
    return scale(_like(distance_map / distance_map.max(), mask))


def _constant_field(mask_image, vector_xyz):
    arr = np.zeros(mask_image.array.shape + (3,))
    arr = arr + np.array([[[vector_xyz]]], dtype=np.float64)
    return _like(arr, mask_image, True)


def _mask_field(field, mask):
    return _like(np.where((mask.array != 0)[..., None], field.array, 0.0), field, True)


def _smooth(field, gaussian_smooth):
    if np.any(gaussian_smooth):
        if not hasattr(gaussian_smooth, "__iter__"):
            gaussian_smooth = (gaussian_smooth,) * 3
        field = _like(orc.recursive_gaussian_vec3(field.array, orc.geom_of(field), gaussian_smooth), field, True)
    return field


def _warp_nn(mask, field):
    tfm = sk.DisplacementFieldTransform(sk.Cast(field, sk.sitkVectorFloat64))
    return ref.apply_transform(mask, transform=tfm, default_value=0, interpolator=sk.sitkNearestNeighbor), tfm


def generate_field_shift(mask_image, vector_shift=(10, 10, 10), gaussian_smooth=5):
    template = _constant_field(mask_image, [-v for v in vector_shift[::-1]])
    shifted, _ = _warp_nn(mask_image, template)
    template = _mask_field(template, _like(mask_image.array | shifted.array, mask_image))
    template = _smooth(template, gaussian_smooth)
    shifted, tfm = _warp_nn(mask_image, template)
    return shifted, tfm, template


def generate_field_asymmetric_contract(mask_image, vector_asymmetric_contract=(10, 10, 10), gaussian_smooth=5, compute_real_dvf=False):
    template = _mask_field(_constant_field(mask_image, list(vector_asymmetric_contract[::-1])), mask_image)
    contracted, _ = _warp_nn(mask_image, template)
    if compute_real_dvf:
        reg_struct = convert_mask_to_reg_structure(mask_image, expansion=3)
        reg_struct_def = convert_mask_to_reg_structure(contracted, expansion=3)
        _, _, template = ref.fast_symmetric_forces_demons_registration(reg_struct_def, reg_struct, isotropic_resample=True,
                                                                       resolution_staging=[4, 2], iteration_staging=[20, 10])
    template = _smooth(template, gaussian_smooth)
    contracted, tfm = _warp_nn(mask_image, template)
    return contracted, tfm, template


def generate_field_asymmetric_extend(mask_image, vector_asymmetric_extend=(10, 10, 10), gaussian_smooth=5):
    template = _constant_field(mask_image, [-v for v in vector_asymmetric_extend[::-1]])
    extended, _ = _warp_nn(mask_image, template)
    template = _smooth(_mask_field(template, extended), gaussian_smooth)
    extended, tfm = _warp_nn(mask_image, template)
    return extended, tfm, template


def generate_field_expand(mask, bone_mask=False, expand=3, gaussian_smooth=5, use_internal_deformation=True):
    mask_original = _like(mask.array + bone_mask.array, mask) if bone_mask is not False else mask
    if not hasattr(expand, "__iter__"):
        expand = (expand,) * 3
    expand = np.array(expand)
    expand = expand / np.array(mask.GetSpacing()[::-1])
    expand = expand[::-1]
    if np.all(np.array(expand) <= 0):
        mask_expand = binary_erode(mask, np.abs(expand).astype(int).tolist())
    elif np.all(np.array(expand) >= 0):
        mask_expand = binary_dilate(mask, np.abs(expand).astype(int).tolist())
    else:
        mask_expand = binary_dilate(mask, np.abs(expand * (expand > 0)).astype(int).tolist())
        mask_expand = binary_erode(mask_expand, np.abs(expand * (expand < 0)).astype(int).tolist())
    if bone_mask is not False:
        mask_expand = _like(mask_expand.array + bone_mask.array, mask)
    if use_internal_deformation:
        reg_original = convert_mask_to_reg_structure(mask_original)
        reg_expand = convert_mask_to_reg_structure(mask_expand)
    else:
        reg_original, reg_expand = mask_original, mask_expand
    _, _, template = ref.fast_symmetric_forces_demons_registration(reg_expand, reg_original, isotropic_resample=True, resolution_staging=[4, 2],
                                                                   iteration_staging=[10, 10], ncores=8)
    template = _smooth(template, gaussian_smooth)
    expanded, tfm = _warp_nn(mask, template)
    return expanded, tfm, template


def generate_field_radial_bend(reference_image, body_mask, reference_point, axis_of_rotation=[0, 0, -1], scale=0.1,
                               mask_bend_from_reference_point=("z", "inf"), gaussian_smooth=5):
    body = body_mask.array.copy()
    where = mask_bend_from_reference_point
    if where is not False:
        if where[0] == "z":
            if where[1] == "inf":
                body[: reference_point[0], :, :] = 0
            elif where[1] == "sup":
                body[reference_point[0]:, :, :] = 0
        if where[0] == "y":
            if where[1] == "post":
                body[:, reference_point[1]:, :] = 0
            elif where[1] == "ant":
                body[:, : reference_point[1], :] = 0
        if where[0] == "x":
            if where[1] == "left":
                body[:, :, reference_point[2]:] = 0
            elif where[1] == "right":
                body[:, :, : reference_point[2]] = 0
    pts = np.array(np.where(body))
    rel = pts - np.array(reference_point)[:, None]
    axis = np.array(axis_of_rotation)
    axis = axis / np.linalg.norm(axis)
    vectors = np.cross(rel[::-1].T, axis[::-1])
    arr = np.zeros(reference_image.array.shape + (3,), dtype=np.float64)
    if scale is not False:
        arr[np.where(body)] = vectors * scale
    template = _smooth(_like(arr, reference_image, True), gaussian_smooth)
    tfm = sk.DisplacementFieldTransform(sk.Cast(template, sk.sitkVectorFloat64))
    bent = ref.apply_transform(reference_image, transform=tfm, default_value=int(reference_image.array.min()), interpolator=sk.sitkLinear)
    return bent, tfm, template


def label_contour(mask, fully_connected=False):
    return _like(orc.label_contour(mask.array, fully_connected), mask)


def evaluate_distance_to_reference(reference_volume, test_volume, resample_factor=1):
    test_distance_map = np.abs(orc.signed_maurer_distance_map(test_volume.array, test_volume.GetSpacing(), False, False, True))
    ref_surface_pts = orc.label_contour(reference_volume.array, False) == 1
    return test_distance_map[ref_surface_pts][::resample_factor]
