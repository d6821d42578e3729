"""
Drop-in replacements for the reference's synthetic-deformation generators and the mask helpers they build on
(consumers that chain the hot path's own pieces -- nearest-neighbour warps, the recursive
Gaussian, Demons -- and therefore stay device resident end to end):

    generate_field_shift                 platipy/imaging/generation/dvf.py:29-81
    generate_field_asymmetric_contract   platipy/imaging/generation/dvf.py:84-156
    generate_field_asymmetric_extend     platipy/imaging/generation/dvf.py:159-216
    generate_field_expand                platipy/imaging/generation/dvf.py:219-324
    generate_field_radial_bend           platipy/imaging/generation/dvf.py:327-415
    convert_mask_to_distance_map         platipy/imaging/registration/utils.py:270-299
    convert_mask_to_reg_structure        platipy/imaging/registration/utils.py:302-344
    ShiftAugment / ExpandAugment / ContractAugment / apply_augmentation / generate_random_augmentation
                                         platipy/imaging/generation/augment.py:33-205
    get_bone_mask                        platipy/imaging/generation/mask.py:21-47

Same arguments, defaults and return values (``(deformed image, DisplacementFieldTransform, displacement field)``).
Vectors follow the reference's convention: given as (z, y, x) in millimetres.  Inputs may be host images or
``DeviceImage`` handles (device in -> device out).
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from collections.abc import Iterable

import numpy as np
import torch

from . import _abi
from . import sitk_compat as sk
from .engine import DeviceImage, Engine
from .label_utils import ball_offsets
from .registration import fast_symmetric_forces_demons_registration

sitkNearestNeighbor, sitkLinear = sk.sitkNearestNeighbor, sk.sitkLinear


def _is_device(image):
    return isinstance(image, DeviceImage)


def _back(eng, dimg, like):
    if _is_device(like):
        eng.release_to_caller()
        return dimg
    return sk.from_native(eng.to_host(dimg), like)


def _transform_of(eng, field):
    """sitk.DisplacementFieldTransform(sitk.Cast(field, sitkVectorFloat64)) around a device-resident field."""
    tfm = sk.DisplacementFieldTransform.__new__(sk.DisplacementFieldTransform)
    tfm._field = field
    tfm._device_cache = (eng, field)
    return tfm


def _finish(eng, image_out, field, like):
    """(image, transform, field) in the caller's representation."""
    tfm = _transform_of(eng, field)
    if _is_device(like):
        eng.release_to_caller()
        return image_out, tfm, field
    host_field = eng.to_host(field)
    tfm._field = host_field
    return sk.from_native(eng.to_host(image_out), like), tfm, sk.from_native(host_field, like)


def _smooth(eng, field, gaussian_smooth):
    # dvf.py:69-74 (and the same block in every generator): sitk.SmoothingRecursiveGaussian(dvf_template, sigma)
    if np.any(gaussian_smooth):
        if not hasattr(gaussian_smooth, "__iter__"):
            gaussian_smooth = (gaussian_smooth,) * 3
        with torch.cuda.stream(eng.stream):
            field = field.like(field.tensor.clone())  # the smoothing entry point works in place
        eng.recursive_gaussian(field, gaussian_smooth)
    return field


def _as_mask(eng, image):
    d = eng.to_device(image)
    if d.is_vector:
        raise RuntimeError("a scalar mask image is expected")
    return eng.cast(d, np.uint8)


def _warp_nn(eng, mask, field):
    # apply_transform(mask, transform=dvf_tfm, default_value=0, interpolator=sitk.sitkNearestNeighbor)
    return eng.resample(mask, mask, _transform_of(eng, field), sitkNearestNeighbor, 0)


# ---------------------------------------------------------------------------------------------------------------------
# registration/utils.py:270-344
# ---------------------------------------------------------------------------------------------------------------------
def _binarise_at_median(eng, mask):
    """utils.py:282-287 / 321-326: a label with more than two distinct positive values is thresholded at their median."""
    lo, hi = eng.minmax(mask)
    if hi <= 2:  # UInt8 values {0, 1, 2}: at most two positive values
        return mask
    with torch.cuda.stream(eng.stream):
        vals = torch.unique(mask.tensor[mask.tensor > 0]).cpu().numpy()  # plumbing: the distinct values of a label image
    if len(vals) > 2:
        cutoff = np.median(vals)
        mask = eng.binary_threshold(mask, float(cutoff), float(np.max(vals)))
    return mask


def _distance_map(eng, mask, squared_distance=False):
    mask = _binarise_at_median(eng, mask)
    return eng.signed_maurer_distance_map(mask, inside_is_positive=True, squared_distance=squared_distance, use_image_spacing=True), mask


def convert_mask_to_distance_map(mask, squared_distance=False, normalise=False):
    """Signed Maurer distance map (positive inside, millimetres) of a binary label (utils.py:270-299) -> Float32."""
    eng = Engine.get()
    raw_map, _ = _distance_map(eng, _as_mask(eng, mask), squared_distance)
    if normalise:
        raw_map = eng.divide_scalar(raw_map, eng.minmax(raw_map)[1])
    return _back(eng, raw_map, mask)


def _reg_structure(eng, mask, expansion=(0, 0, 0)):
    mask = _binarise_at_median(eng, mask)
    if not hasattr(expansion, "__iter__"):
        expansion = [int(expansion / i) for i in mask.GetSpacing()]
    if any(expansion):
        mask = eng.binary_dilate(mask, ball_offsets(expansion))  # sitk.BinaryDilate(mask, expansion): ball kernel
    distance_map, mask = _distance_map(eng, mask, False)
    distance_map = eng.mask_image(eng.cast(distance_map, np.float64), mask)
    return eng.divide_scalar(distance_map, eng.minmax(distance_map)[1])


def convert_mask_to_reg_structure(mask, expansion=(0, 0, 0), scale=lambda x: x):
    """A mask-like Float64 image (distance to the surface inside the label, scaled to [0, 1]) that makes structure-guided
    registration deform the interior as well (utils.py:302-344).  ``scale`` is applied to the result as in the reference; it
    receives the image in the caller's representation."""
    eng = Engine.get()
    out = _reg_structure(eng, _as_mask(eng, mask), expansion)
    return scale(_back(eng, out, mask))


# ---------------------------------------------------------------------------------------------------------------------
# generation/dvf.py
# ---------------------------------------------------------------------------------------------------------------------
def generate_field_shift(mask_image, vector_shift=(10, 10, 10), gaussian_smooth=5):
    """Shifts a structure defined by a binary mask (dvf.py:29-81)."""
    eng = Engine.get()
    mask = _as_mask(eng, mask_image)
    shift_xyz = [-float(v) for v in vector_shift[::-1]]
    # the constant field everywhere, warp, then keep it inside (mask | shifted mask)
    shifted = _warp_nn(eng, mask, eng.constant_field(mask, shift_xyz))
    field = eng.constant_field(mask, shift_xyz, eng.u8_binary_op(mask, shifted, _abi.OP_OR))
    field = _smooth(eng, field, gaussian_smooth)
    return _finish(eng, _warp_nn(eng, mask, field), field, mask_image)


def generate_field_asymmetric_contract(mask_image, vector_asymmetric_contract=(10, 10, 10), gaussian_smooth=5, compute_real_dvf=False):
    """Contracts a structure along a vector (dvf.py:84-156)."""
    eng = Engine.get()
    mask = _as_mask(eng, mask_image)
    field = eng.constant_field(mask, [float(v) for v in vector_asymmetric_contract[::-1]], mask)
    if compute_real_dvf:
        contracted = _warp_nn(eng, mask, field)
        reg_struct = _reg_structure(eng, mask, expansion=3)
        reg_struct_def = _reg_structure(eng, contracted, expansion=3)
        _, _, field = fast_symmetric_forces_demons_registration(reg_struct_def, reg_struct, isotropic_resample=True, resolution_staging=[4, 2],
                                                                iteration_staging=[20, 10])
        eng.wait_caller()
    field = _smooth(eng, field, gaussian_smooth)
    return _finish(eng, _warp_nn(eng, mask, field), field, mask_image)


def generate_field_asymmetric_extend(mask_image, vector_asymmetric_extend=(10, 10, 10), gaussian_smooth=5):
    """Extends a structure along a vector (dvf.py:159-216)."""
    eng = Engine.get()
    mask = _as_mask(eng, mask_image)
    extend_xyz = [-float(v) for v in vector_asymmetric_extend[::-1]]
    extended = _warp_nn(eng, mask, eng.constant_field(mask, extend_xyz))
    field = _smooth(eng, eng.constant_field(mask, extend_xyz, extended), gaussian_smooth)
    return _finish(eng, _warp_nn(eng, mask, field), field, mask_image)


def generate_field_expand(mask, bone_mask=False, expand=3, gaussian_smooth=5, use_internal_deformation=True):
    """Expands / shrinks a structure with a ball whose radii are ``expand`` (z, y, x; millimetres), then finds the
    deformation between the two with Demons (dvf.py:219-324)."""
    eng = Engine.get()
    m = _as_mask(eng, mask)
    bone = _as_mask(eng, bone_mask) if bone_mask is not False else None
    mask_original = eng.u8_binary_op(m, bone, _abi.OP_ADD) if bone is not None else m

    if not hasattr(expand, "__iter__"):
        expand = (expand,) * 3
    expand = np.array(expand)
    expand = expand / np.array(m.GetSpacing()[::-1])  # dvf.py:260
    expand = expand[::-1]  # (x, y, z)

    def radius(v):
        return np.abs(v).astype(int).tolist()

    if np.all(expand <= 0):
        print("All factors negative: shrinking only.")
        mask_expand = eng.binary_erode(m, ball_offsets(radius(expand)))
    elif np.all(expand >= 0):
        print("All factors positive: expansion only.")
        mask_expand = eng.binary_dilate(m, ball_offsets(radius(expand)))
    else:
        print("Mixed factors: shrinking and expansion.")
        mask_expand = eng.binary_dilate(m, ball_offsets(radius(expand * (expand > 0))))
        mask_expand = eng.binary_erode(mask_expand, ball_offsets(radius(expand * (expand < 0))))
    if bone is not None:
        mask_expand = eng.u8_binary_op(mask_expand, bone, _abi.OP_ADD)

    if use_internal_deformation:
        registration_mask_original = _reg_structure(eng, mask_original)
        registration_mask_expand = _reg_structure(eng, mask_expand)
    else:
        registration_mask_original, registration_mask_expand = mask_original, mask_expand

    _, _, field = fast_symmetric_forces_demons_registration(registration_mask_expand, registration_mask_original, isotropic_resample=True,
                                                            resolution_staging=[4, 2], iteration_staging=[10, 10], ncores=8)
    eng.wait_caller()
    field = _smooth(eng, field, gaussian_smooth)
    return _finish(eng, _warp_nn(eng, m, field), field, mask)


_BEND_CLIP = {  # (dimension, limit) -> (axis in x/y/z order, keep voxels with index >= reference); dvf.py:364-379
    ("z", "inf"): (2, True), ("z", "sup"): (2, False),
    ("y", "post"): (1, False), ("y", "ant"): (1, True),
    ("x", "left"): (0, False), ("x", "right"): (0, True),
}


def generate_field_radial_bend(reference_image, body_mask, reference_point, axis_of_rotation=[0, 0, -1], scale=0.1,
                               mask_bend_from_reference_point=("z", "inf"), gaussian_smooth=5):
    """A field of rotation-like bending about ``reference_point`` (z, y, x voxel index) inside the body mask
    (dvf.py:327-415): displacement = scale * (voxel - reference_point) x axis_of_rotation."""
    eng = Engine.get()
    image = eng.to_device(reference_image)
    body = _as_mask(eng, body_mask)
    size_xyz = body.GetSize()
    ref_xyz = [int(v) for v in reference_point][::-1]
    # numpy slice semantics of body_mask_arr[: r] / [r :] for a negative r
    ref_xyz = [r if r >= 0 else max(0, n + r) for r, n in zip(ref_xyz, size_xyz)]
    clip_axis, keep_upper = -1, True
    if mask_bend_from_reference_point is not False:
        clip_axis, keep_upper = _BEND_CLIP.get((mask_bend_from_reference_point[0], mask_bend_from_reference_point[1]), (-1, True))
    axis = np.array(axis_of_rotation)
    axis = axis / np.linalg.norm(axis)
    if scale is not False:
        ref_for_vectors = [int(v) for v in reference_point][::-1]
        if ref_for_vectors != ref_xyz:
            raise NotImplementedError("generate_field_radial_bend: negative reference_point indices are not supported")
        field = eng.radial_bend_field(body, ref_xyz, axis[::-1], scale, clip_axis, keep_upper)
    else:
        field = eng.constant_field(body, (0.0, 0.0, 0.0))
    if field.GetSize() != image.GetSize():
        raise RuntimeError("CopyInformation: the body mask and the reference image must have the same size")  # dvf.py:397
    field = DeviceImage(field.tensor, np.float64, image.GetSpacing(), image.GetOrigin(), image.GetDirection(), True)
    field = _smooth(eng, field, gaussian_smooth)
    default_value = int(eng.minmax(image)[0])  # dvf.py:411
    bent = eng.resample(image, image, _transform_of(eng, field), sitkLinear, default_value)
    return _finish(eng, bent, field, reference_image)


# ---------------------------------------------------------------------------------------------------------------------
# generation/mask.py
# ---------------------------------------------------------------------------------------------------------------------
def get_bone_mask(image, lower_threshold=350, upper_threshold=3500, max_hole_size=5):
    """Binary mask of the bones of a CT image (mask.py:21-47): BinaryThreshold, then BinaryMorphologicalClosing with
    ``max_hole_size`` handed to SimpleITK as the kernel radius (voxels, (x, y, z) for a vector -- the reference passes it
    through unchanged although its docstring speaks of millimetres and (z, y, x))."""
    from .label_utils import binary_morphological_closing

    eng = Engine.get()
    d = eng.to_device(image)
    bone = eng.binary_threshold(d, float(lower_threshold), float(upper_threshold))
    if max_hole_size is False:
        # mask.py:41-45: the closing runs in any case; SimpleITK rejects a bool radius
        raise TypeError("BinaryMorphologicalClosing: kernelRadius must be an int or a sequence of ints, not bool")
    if not hasattr(max_hole_size, "__iter__"):
        max_hole_size = (max_hole_size,) * 3
    closed = binary_morphological_closing(bone, max_hole_size)
    eng.wait_caller()
    return _back(eng, closed, image)


# ---------------------------------------------------------------------------------------------------------------------
# generation/augment.py
# ---------------------------------------------------------------------------------------------------------------------
class DeformableAugment(ABC):
    @abstractmethod
    def augment(self):
        """-> (transform, displacement field)"""


class ShiftAugment(DeformableAugment):
    def __init__(self, mask, vector_shift=(10, 10, 10), gaussian_smooth=5):
        self.mask, self.vector_shift, self.gaussian_smooth = mask, vector_shift, gaussian_smooth

    def augment(self):
        _, transform, dvf = generate_field_shift(self.mask, self.vector_shift, self.gaussian_smooth)
        return transform, dvf


class ExpandAugment(DeformableAugment):
    def __init__(self, mask, vector_expand=(10, 10, 10), gaussian_smooth=5, bone_mask=False):
        self.mask, self.vector_expand, self.gaussian_smooth, self.bone_mask = mask, vector_expand, gaussian_smooth, bone_mask

    def augment(self):
        _, transform, dvf = generate_field_expand(self.mask, bone_mask=self.bone_mask, expand=self.vector_expand,
                                                  gaussian_smooth=self.gaussian_smooth)
        return transform, dvf


class ContractAugment(DeformableAugment):
    def __init__(self, mask, vector_contract=(10, 10, 10), gaussian_smooth=5, bone_mask=False):
        self.mask = mask
        self.contract = [int(-x / s) for x, s in zip(vector_contract, mask.GetSpacing())]  # augment.py:193
        self.gaussian_smooth, self.bone_mask = gaussian_smooth, bone_mask

    def augment(self):
        _, transform, dvf = generate_field_expand(self.mask, bone_mask=self.bone_mask, expand=self.contract, gaussian_smooth=self.gaussian_smooth)
        return transform, dvf


def apply_augmentation(image, augmentation, masks=[]):
    """Apply one or several augmentations to an image and its masks (augment.py:33-83): the transforms are composed, the
    fields summed; returns ``(image, [masks,] field)``."""
    if not (isinstance(image, (sk.Image, DeviceImage)) or sk.is_native_sitk(image)):
        raise AttributeError("image should be a SimpleITK.Image")
    if isinstance(augmentation, DeformableAugment):
        augmentation = [augmentation]
    if not isinstance(augmentation, Iterable):
        raise AttributeError("augmentation must be a DeformableAugment or an iterable (such as list) of DeformableAugment's")
    eng = Engine.get()
    transforms, dvf = [], None
    for aug in augmentation:
        if not isinstance(aug, DeformableAugment):
            raise AttributeError("Each augmentation must be of type DeformableAugment")
        tfm, field = aug.augment()
        transforms.append(tfm)
        f = eng.to_device(field)
        if dvf is None:
            dvf = f
        else:
            with torch.cuda.stream(eng.stream):  # dvf += field: plumbing-level add of two device fields
                dvf = dvf.like(dvf.tensor + f.tensor)
    if len(transforms) > _abi.MAX_TRANSFORMS:
        raise ValueError(f"at most {_abi.MAX_TRANSFORMS} augmentations can be composed")
    transform = sk.CompositeTransform(transforms)
    d = eng.to_device(image)
    image_deformed = _back(eng, eng.resample(d, d, transform, sitkLinear, int(eng.minmax(d)[0])), image)
    masks_deformed = []
    for mask in masks:
        dm = eng.to_device(mask)
        masks_deformed.append(_back(eng, eng.resample(dm, dm, transform, sitkNearestNeighbor, 0), mask))
    dvf_out = _back(eng, dvf, image)
    if masks:
        return image_deformed, masks_deformed, dvf_out
    return image_deformed, dvf_out


def generate_random_augmentation(ct_image, masks):
    """One randomly chosen and randomly parameterised augmentation per mask (augment.py:86-141): shifts of up to 10 mm,
    contractions and expansions of up to 10 mm with the bone mask held fixed, field smoothing of 3-5 mm."""
    import random

    random.shuffle(masks)
    augmentation_types = [
        {"class": ShiftAugment, "args": {"vector_shift": [(-10, 10), (10, 10), (-10, 10)], "gaussian_smooth": (3, 5)}},
        {"class": ContractAugment, "args": {"vector_contract": [(0, 10), (0, 10), (0, 10)], "gaussian_smooth": (3, 5), "bone_mask": True}},
        {"class": ExpandAugment, "args": {"vector_expand": [(0, 10), (0, 10), (0, 10)], "gaussian_smooth": (3, 5), "bone_mask": True}},
    ]
    augmentation = []
    for mask in masks:
        aug = random.choice(augmentation_types)
        aug_args = {}
        for arg, spec in aug["args"].items():
            value = spec
            if isinstance(spec, list):  # one draw per dimension
                value = [random.randint(lo, hi) for lo, hi in spec]
            elif isinstance(spec, tuple):
                value = random.randint(spec[0], spec[1])
            if arg == "bone_mask" and spec:
                value = get_bone_mask(ct_image)
            aug_args[arg] = value
        augmentation.append(aug["class"](mask, **aug_args))
    return augmentation
