"""
CPU restatement of the reference's label comparison metrics (platipy/imaging/label/comparison.py) -- TEST INFRASTRUCTURE, NOT
PRODUCT CODE.

PINNED: unlike the rest of the oracle, these functions are checked against outputs of the real SimpleITK: the reference's own
known-answer tests (platipy/imaging/tests/test_metrics.py:6-67) hold eleven golden numbers for ``compute_surface_dsc`` and
``compute_surface_metrics``, and ``tests/test_reference_golden_metrics.py`` reproduces every one of them (the means to the
last digit).  That pins, against ITK itself, the restatements these metrics are made of:

  sitk.SignedMaurerDistanceMap             itk_oracle.c ``orc_signed_maurer`` (single-precision Voronoi passes, distance to the
                                           26-neighbourhood contour, sign, image spacing)
  sitk.BinaryContourImageFilter (FullyConnectedOn), sitk.LabelContour (face neighbours)      ``orc_label_contour``
  sitk.HausdorffDistanceImageFilter        max over the object of max(distance map of the other object, 0), both directions
  sitk.LabelIntensityStatisticsImageFilter mean, maximum, unbiased standard deviation, and the median as the centre of the
                                           histogram bin (128 bins over the feature image's global range) where the cumulative
                                           count reaches half

Function names and argument meaning follow the reference.  Images are ``platipy_b200.sitk_compat.Image`` containers.
"""
from __future__ import annotations

import numpy as np

from platipy_b200.sitk_compat import Image

from . import itk_oracle as orc
from . import platipy_ref as ref


def _like(arr, like):
    return Image(arr, like.GetSpacing(), like.GetOrigin(), like.GetDirection())


def signed_maurer(label, **kw):
    return orc.signed_maurer_distance_map(label.array, label.GetSpacing(), **kw)


def label_intensity_statistics(label_arr, feature, n_bins=128):
    """itk::LabelIntensityStatisticsImageFilter for label 1: (mean, maximum, standard deviation, median, number of pixels)."""
    vals = feature[label_arr == 1].astype(np.float64)
    fmin, fmax = float(feature.min()), float(feature.max())
    width = (fmax - fmin) / n_bins
    idx = np.minimum(np.floor((vals - fmin) / width).astype(np.int64), n_bins - 1)
    cum = np.cumsum(np.bincount(idx, minlength=n_bins))
    median = fmin + (int(np.argmax(cum >= vals.size / 2)) + 0.5) * width
    return vals.mean(), vals.max(), vals.std(ddof=1), median, vals.size


def hausdorff_distance(label_a, label_b):
    """itk::HausdorffDistanceImageFilter (UseImageSpacing on): the larger of the two directed distances.  ITK's directed filter works
    on a distance map in ITS real type (double); the Float32 restatement is therefore asked for SQUARED distances -- exact sums of
    squares for spacings like the reference's test (1, 1, 2) -- and the root is taken in double: sqrt(6) and sqrt(150) come out as
    the very doubles the reference's test_metrics.py lists."""
    out = []
    for la, lb in ((label_a, label_b), (label_b, label_a)):
        d2 = signed_maurer(lb, squared_distance=True)
        out.append(float(np.sqrt(np.float64(np.maximum(d2[la.array != 0], 0).max()))))
    return max(out)


def compute_volume(label):
    return label.array.sum() * np.prod(label.GetSpacing()) / 1000


def compute_surface_dsc(label_a, label_b, tau=3.0):
    # comparison.py:35-72
    a_contour, b_contour = orc.label_contour(label_a.array, True), orc.label_contour(label_b.array, True)
    dist_to_a = orc.signed_maurer_distance_map(a_contour, label_a.GetSpacing())
    dist_to_b = orc.signed_maurer_distance_map(b_contour, label_b.GetSpacing())
    b_intersection = (b_contour * (dist_to_a <= tau)).sum()
    a_intersection = (a_contour * (dist_to_b <= tau)).sum()
    return (b_intersection + a_intersection) / (a_contour.sum() + b_contour.sum())


def _surface_statistics(label_a, label_b):
    rows = []
    for la, lb in ((label_a, label_b), (label_b, label_a)):
        reference_distance_map = np.abs(signed_maurer(la))
        rows.append(label_intensity_statistics(orc.label_contour(lb.array, False), reference_distance_map))
    return rows


def compute_surface_metrics(label_a, label_b, verbose=False):
    # comparison.py:75-141
    rows = _surface_statistics(label_a, label_b)
    mean_sd_list, max_sd_list, std_sd_list, median_sd_list, num_points = (list(c) for c in zip(*rows))
    mean_surf_dist = np.dot(mean_sd_list, num_points) / np.sum(num_points)
    return {
        "hausdorffDistance": hausdorff_distance(label_a, label_b),
        "hausdorffDistance95": np.percentile(max_sd_list, 95),
        "meanSurfaceDistance": mean_surf_dist,
        "medianSurfaceDistance": np.mean(median_sd_list),
        "maximumSurfaceDistance": np.max(max_sd_list),
        "sigmaSurfaceDistance": np.sqrt(np.dot(num_points, np.add(np.square(std_sd_list), np.square(np.subtract(mean_sd_list, mean_surf_dist))))),
        "surfaceDSC": compute_surface_dsc(label_a, label_b),
    }


def compute_volume_metrics(label_a, label_b):
    # comparison.py:144-191
    a, b = label_a.array.astype(bool), label_b.array.astype(bool)
    inter, union = a & b, a | b
    voxel_volume = np.prod(label_a.GetSpacing()) / 1000.0
    true_pos, true_neg = inter.sum(), (~a & ~b).sum()
    false_pos, false_neg = b.sum() - true_pos, a.sum() - true_pos
    return {
        "DSC": (2.0 * inter.sum()) / (a.sum() + b.sum()),
        "volumeOverlap": inter.sum() * voxel_volume,
        "fractionOverlap": inter.sum() / union.sum().astype(float),
        "truePositiveFraction": (1.0 * true_pos) / (true_pos + false_neg),
        "trueNegativeFraction": (1.0 * true_neg) / (true_neg + false_pos),
        "falsePositiveFraction": (1.0 * false_pos) / (true_neg + false_pos),
        "falseNegativeFraction": (1.0 * false_neg) / (true_pos + false_neg),
    }


def _auto_crop(label_a, label_b):
    # comparison.py:205-210: crop both labels to the bounding box of their union
    union = _like(((label_a.array.astype(np.int64) + label_b.array) > 0).astype(np.uint8), label_a)
    size, index = ref.label_to_roi(union)
    return ref.crop_to_roi(label_a, size, index), ref.crop_to_roi(label_b, size, index)


def compute_metric_dsc(label_a, label_b, auto_crop=True):
    if auto_crop:
        label_a, label_b = _auto_crop(label_a, label_b)
    a, b = label_a.array.astype(bool), label_b.array.astype(bool)
    return 2 * ((a & b).sum()) / (a.sum() + b.sum())


def compute_metric_specificity(label_a, label_b, auto_crop=True):
    if auto_crop:
        label_a, label_b = _auto_crop(label_a, label_b)
    a, b = label_a.array.astype(bool), label_b.array.astype(bool)
    true_pos, true_neg = (a & b).sum(), (~a & ~b).sum()
    return float((1.0 * true_neg) / (true_neg + (b.sum() - true_pos)))


def compute_metric_sensitivity(label_a, label_b, auto_crop=True):
    if auto_crop:
        label_a, label_b = _auto_crop(label_a, label_b)
    a, b = label_a.array.astype(bool), label_b.array.astype(bool)
    true_pos = (a & b).sum()
    return float((1.0 * true_pos) / (true_pos + (a.sum() - true_pos)))


def compute_metric_masd(label_a, label_b, auto_crop=True):
    if auto_crop:
        label_a, label_b = _auto_crop(label_a, label_b)
    if label_a.array.sum() == 0 or label_b.array.sum() == 0:
        return np.nan
    rows = _surface_statistics(label_a, label_b)
    means, counts = [r[0] for r in rows], [r[4] for r in rows]
    return float(np.dot(means, counts) / np.sum(counts))


def compute_metric_hd(label_a, label_b, auto_crop=True):
    if auto_crop:
        label_a, label_b = _auto_crop(label_a, label_b)
    if label_a.array.sum() == 0 or label_b.array.sum() == 0:
        return np.nan
    return hausdorff_distance(label_a, label_b)


def compute_apl(label_ref, label_test, distance_threshold_mm=3):
    # comparison.py:346-387, slice by slice like the reference; the 2-D LabelContour / BinaryDilate are the 3-D restatements on a one-slice volume
    from .generation_ref import ball

    added = []
    distance = int(np.ceil(distance_threshold_mm / np.mean(label_ref.GetSpacing()[:2])))
    for i in range(label_ref.array.shape[0]):
        ref_slice, test_slice = label_ref.array[i:i + 1], label_test.array[i:i + 1]
        if int(ref_slice.sum()) + int(test_slice.sum()) == 0:
            continue
        ref_contour, test_contour = orc.label_contour(ref_slice, False), orc.label_contour(test_slice, False)
        if distance_threshold_mm > 0:
            test_contour = orc.binary_morph(test_contour, ball([distance, distance, 0]), True)
        added.append(np.where(test_contour == 0, ref_contour, 0).sum())
    return added


def compute_metric_total_apl(label_ref, label_test, distance_threshold_mm=3):
    return np.sum(compute_apl(label_ref, label_test, distance_threshold_mm)) * np.mean(label_ref.GetSpacing()[:2])


def compute_metric_mean_apl(label_ref, label_test, distance_threshold_mm=3):
    return np.mean(compute_apl(label_ref, label_test, distance_threshold_mm)) * np.mean(label_ref.GetSpacing()[:2])
