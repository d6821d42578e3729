"""
Drop-in replacements for the label utilities the atlas pipeline calls around the fusion step:

    label_to_roi, crop_to_roi, crop_to_label_extent   platipy/imaging/utils/crop.py:24-99
    correct_volume_overlap                            platipy/imaging/label/utils.py:23-58
    paste                                             sitk.Paste as used in multiatlas/run.py:387-404
    binary_morphological_closing                      sitk.BinaryMorphologicalClosing as used in multiatlas/run.py:424
    get_com                                           platipy/imaging/label/utils.py:61-84

Same arguments and return values; inputs may be host images or ``DeviceImage`` handles (device in -> device out).
"""
from __future__ import annotations

import numpy as np
import torch

from . import sitk_compat as sk
from .engine import DeviceImage, Engine


def _back(eng, dimg, like):
    if isinstance(like, DeviceImage):
        eng.release_to_caller()
        return dimg
    return sk.from_native(eng.to_host(dimg), like)


def label_to_roi(label, expansion_mm=[0, 0, 0], return_as_list=False):
    """Region of interest (size, index) of a label or a list of labels, expanded by ``expansion_mm`` per direction and
    clipped to the image (crop.py:24-71).  Index / size are in (x, y, z) order like SimpleITK's."""
    eng = Engine.get()
    if hasattr(label, "__iter__") and not isinstance(label, (sk.Image, DeviceImage)) and not sk.is_native_sitk(label):
        labels = [eng.to_device(l) for l in label]
        mask = eng.binary_threshold(labels[0], 1e-300, np.inf)
        for l in labels[1:]:  # sum(label) > 0
            m2 = eng.binary_threshold(l, 1e-300, np.inf)
            with torch.cuda.stream(eng.stream):  # torch is plumbing here: an OR of two device masks, on the engine's stream
                mask = mask.like(mask.tensor | m2.tensor, np.uint8, False)
    else:
        d = eng.to_device(label)
        mask = eng.binary_threshold(d, 1e-300, np.inf)  # label > 0
    bb = eng.bounding_box(mask)
    if bb[3] < bb[0]:
        raise RuntimeError("label_to_roi: the label is empty (LabelStatisticsImageFilter has no label 1)")
    spacing = np.array(mask.GetSpacing())
    index = np.array(bb[:3])
    size = np.array([bb[3 + k] - bb[k] + 1 for k in range(3)])
    expansion = (np.array(expansion_mm) / spacing).astype(int)
    crop_box_index = np.max([index - expansion, np.array([0, 0, 0])], axis=0)
    crop_box_size = np.min([np.array(mask.GetSize()) - crop_box_index, size + 2 * expansion], axis=0)
    crop_box_size = [int(i) for i in crop_box_size]
    crop_box_index = [int(i) for i in crop_box_index]
    if return_as_list:
        return crop_box_index + crop_box_size
    return crop_box_size, crop_box_index


def crop_to_roi(image, size, index):
    """sitk.RegionOfInterest(image, size=size, index=index) (crop.py:74-76): the origin moves to the first voxel kept."""
    eng = Engine.get()
    d = eng.to_device(image)
    if d.is_vector:
        raise NotImplementedError("crop_to_roi of vector images is not implemented")
    sx, sy, sz = (int(v) for v in size)
    out = eng.empty((sz, sy, sx), d.np_dtype)
    direction = np.asarray(d.GetDirection(), dtype=np.float64).reshape(3, 3)
    origin = np.asarray(d.GetOrigin()) + direction @ (np.asarray(d.GetSpacing()) * np.asarray(index, dtype=np.float64))
    res = DeviceImage(out, d.np_dtype, d.GetSpacing(), tuple(origin), d.GetDirection(), False)
    eng.region_copy(d, index, res, (0, 0, 0), (sx, sy, sz))
    return _back(eng, res, image)


def crop_to_label_extent(image, label, expansion_mm=0):
    """crop.py:79-99 (note: the reference's ``~hasattr`` test is always true, so ``expansion_mm`` is always replicated)."""
    expansion_mm = [expansion_mm] * 3
    size, index = label_to_roi(label, expansion_mm=expansion_mm)
    return crop_to_roi(image, size, index)


def paste(destination_image, source_image, source_size=None, source_index=(0, 0, 0), destination_index=(0, 0, 0)):
    """sitk.Paste(destination, source, sourceSize, sourceIndex, destinationIndex): a copy of ``destination_image`` with the
    source region written at ``destination_index`` (run.py:387-404); pixel types must agree."""
    eng = Engine.get()
    dst, src = eng.to_device(destination_image), eng.to_device(source_image)
    if dst.np_dtype != src.np_dtype:
        raise RuntimeError("Paste: both images must have the same pixel type")
    if source_size is None:
        source_size = src.GetSize()
    with torch.cuda.stream(eng.stream):
        out = dst.like(dst.tensor.clone())
    eng.region_copy(src, source_index, out, destination_index, source_size)
    return _back(eng, out, destination_image)


def correct_volume_overlap(binary_label_dict, assign_overlap_to_largest=True):
    """label/utils.py:23-58: structures are ranked by volume (largest first by default) and every voxel is kept by the
    first structure in that order that contains it."""
    eng = Engine.get()
    keys = list(binary_label_dict.keys())
    dev = {k: eng.cast(eng.to_device(binary_label_dict[k]), np.uint8) for k in keys}
    with torch.cuda.stream(eng.stream):
        vals = [int(dev[k].tensor.sum(dtype=torch.int64).item()) for k in keys]  # .sum() of the array view
    volume_rank = np.argsort(vals)[::-1] if assign_overlap_to_largest else np.argsort(vals)
    ranked = [keys[i] for i in volume_rank]
    binar = [eng.binary_threshold(dev[k], 1e-300, np.inf) for k in ranked]  # s_img > 0
    outs = eng.resolve_overlap(binar)
    return {k: _back(eng, o, binary_label_dict[k]) for k, o in zip(ranked, outs)}


def get_com(label, as_int=True, real_coords=False):
    """Centre of mass of a label image (label/utils.py:61-84; scipy.ndimage.center_of_mass of the array): array order
    (z, y, x) -- truncated to ints by default --, or the physical point (x, y, z) with ``real_coords``."""
    eng = Engine.get()
    d = eng.cast(eng.to_device(label), np.float32)
    if d.is_vector:
        raise RuntimeError("a scalar image is expected")
    # first moments in INDEX space: the moments kernel on the same pixels with an identity geometry
    on_index_grid = DeviceImage(d.tensor, np.float32, (1.0, 1.0, 1.0), (0.0, 0.0, 0.0), (1, 0, 0, 0, 1, 0, 0, 0, 1), False)
    m = eng.image_moments(on_index_grid)
    with np.errstate(divide="ignore", invalid="ignore"):
        com_xyz = m[1:4] / m[0]  # scipy returns nan for an all-zero image (with a warning), and so does this
    if real_coords:
        direction = np.asarray(d.GetDirection(), dtype=np.float64).reshape(3, 3)
        return tuple(np.asarray(d.GetOrigin()) + direction @ (np.asarray(d.GetSpacing()) * com_xyz))  # TransformContinuousIndexToPhysicalPoint
    com = tuple(float(v) for v in com_xyz[::-1])
    if as_int:
        return [int(i) for i in com]
    return com


def ball_offsets(radius):
    """Offsets (dx, dy, dz) of itk::FlatStructuringElement::Ball(radius) (SimpleITK's default sitkBall kernel): the voxels of the
    (2r+1)^3 box whose centre lies inside the ellipsoid with semi-axes r + 1/2.  [ITK-recall: the exact inclusion rule of
    ITK's ellipsoid flood fill could not be re-read here; this is the symmetric form.]"""
    r = [int(v) for v in radius]
    ax = [rr + 0.5 for rr in r]
    offs = []
    for dz in range(-r[2], r[2] + 1):
        for dy in range(-r[1], r[1] + 1):
            for dx in range(-r[0], r[0] + 1):
                if (dx / ax[0]) ** 2 + (dy / ax[1]) ** 2 + (dz / ax[2]) ** 2 <= 1.0:
                    offs.append((dx, dy, dz))
    return np.array(offs, dtype=np.int32)


def binary_morphological_closing(image, kernel_radius=(1, 1, 1)):
    """sitk.BinaryMorphologicalClosing(image, kernelRadius) with the default ball kernel, foreground 1, SafeBorder on."""
    eng = Engine.get()
    d = eng.cast(eng.to_device(image), np.uint8)
    radius = [int(kernel_radius)] * 3 if np.isscalar(kernel_radius) else [int(v) for v in kernel_radius]
    out = eng.binary_closing(d, radius, ball_offsets(radius))
    return _back(eng, out, image)
