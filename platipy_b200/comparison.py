"""
Drop-in replacements for the reference's label comparison metrics (platipy/imaging/label/comparison.py) -- the functions that
score what the atlas pipeline produces:

    compute_volume, compute_volume_metrics, compute_metric_dsc / _specificity / _sensitivity      comparison.py:22-270
    compute_surface_dsc, compute_surface_metrics, compute_metric_masd, compute_metric_hd          comparison.py:35-343

The volume work (distance maps, contours, masking, bounding boxes, crops) runs on the device; what comes back is a handful of
counts, extrema and the distances AT the contour voxels (a few thousand numbers per label), on which the statistics of
itk::LabelIntensityStatisticsImageFilter are formed exactly as ITK forms them (float64 mean / unbiased standard deviation of the
Float32 distances, median = centre of the histogram bin -- 128 bins over the distance map's global range -- where the cumulative
count reaches half).  These are the functions the reference has known-answer tests for (platipy/imaging/tests/test_metrics.py);
tests/test_gpu_zzz_session3.py asserts the same golden numbers.

    compute_apl, compute_metric_total_apl, compute_metric_mean_apl                                comparison.py:346-431
"""
from __future__ import annotations

import numpy as np
import torch

from . import _abi
from . import label_utils as lu
from .engine import Engine

_HISTOGRAM_BINS = 128  # itk::StatisticsLabelMapFilter's default for pixel types wider than one byte


def _mask(eng, image):
    d = eng.to_device(image)
    if d.is_vector:
        raise RuntimeError("a scalar label image is expected")
    return eng.cast(d, np.uint8)


def _count(eng, mask):
    with torch.cuda.stream(eng.stream):  # plumbing: number of non-zero voxels of a device mask
        return int(torch.count_nonzero(mask.tensor).item())


def _values_at(eng, image, where):
    """image[where == 1] in raster order, on the host (plumbing: stream compaction of a few thousand values)."""
    with torch.cuda.stream(eng.stream):
        host = image.tensor[where.tensor == 1].cpu()
    eng.synchronize()
    return host.numpy()


def _auto_crop(eng, a, b):
    # comparison.py:205-210: largest_region = (label_a + label_b) > 0; crop both labels to its bounding box
    union = eng.u8_binary_op(eng.binary_threshold(a, 1e-300, np.inf), eng.binary_threshold(b, 1e-300, np.inf), _abi.OP_OR)
    size, index = lu.label_to_roi(union)
    return lu.crop_to_roi(a, size, index), lu.crop_to_roi(b, size, index)


def _pair(label_a, label_b, auto_crop):
    eng = Engine.get()
    a, b = _mask(eng, label_a), _mask(eng, label_b)
    if a.GetSize() != b.GetSize():
        raise RuntimeError("the two labels do not occupy the same grid")
    if auto_crop:
        a, b = _auto_crop(eng, a, b)
        eng.wait_caller()
    return eng, a, b


def _overlap_counts(eng, a, b):
    """(|A|, |B|, |A and B|, number of voxels) with A, B the non-zero voxels."""
    ba, bb = eng.binary_threshold(a, 1e-300, np.inf), eng.binary_threshold(b, 1e-300, np.inf)
    # numpy scalars: a zero denominator gives nan / inf with numpy's warning, as the reference's array arithmetic does, not ZeroDivisionError
    return (np.float64(_count(eng, ba)), np.float64(_count(eng, bb)), np.float64(_count(eng, eng.u8_binary_op(ba, bb, _abi.OP_AND))),
            np.float64(a.GetNumberOfPixels()))


# ---------------------------------------------------------------------------------------------------------------------
# volume metrics
# ---------------------------------------------------------------------------------------------------------------------
def compute_volume(label):
    """Volume in cubic centimetres (comparison.py:22-32): the SUM of the label's values times the voxel volume."""
    eng = Engine.get()
    d = eng.to_device(label)
    with torch.cuda.stream(eng.stream):  # plumbing: the sum of a label image
        total = d.tensor.sum(dtype=torch.float64 if d.np_dtype.kind == "f" else torch.int64).item()
    return total * np.prod(d.GetSpacing()) / 1000


def compute_volume_metrics(label_a, label_b):
    """DSC, volumeOverlap, fractionOverlap and the true / false positive / negative fractions (comparison.py:144-191)."""
    eng, a, b = _pair(label_a, label_b, False)
    na, nb, inter, n = _overlap_counts(eng, a, b)
    union = na + nb - inter
    voxel_volume = np.prod(a.GetSpacing()) / 1000.0
    true_pos, true_neg = inter, n - union
    false_pos, false_neg = nb - true_pos, na - true_pos
    return {
        "DSC": (2.0 * inter) / (na + nb),
        "volumeOverlap": inter * voxel_volume,
        "fractionOverlap": inter / union,
        "truePositiveFraction": (1.0 * true_pos) / (true_pos + false_neg),
        "trueNegativeFraction": (1.0 * true_neg) / (true_neg + false_pos),
        "falsePositiveFraction": (1.0 * false_pos) / (true_neg + false_pos),
        "falseNegativeFraction": (1.0 * false_neg) / (true_pos + false_neg),
    }


def compute_metric_dsc(label_a, label_b, auto_crop=True):
    """Dice similarity coefficient (comparison.py:194-213)."""
    eng, a, b = _pair(label_a, label_b, auto_crop)
    na, nb, inter, _ = _overlap_counts(eng, a, b)
    return 2 * inter / (na + nb)


def compute_metric_specificity(label_a, label_b, auto_crop=True):
    """comparison.py:216-242 (the true negatives are counted inside the cropped region when auto_crop is on, as in the reference)."""
    eng, a, b = _pair(label_a, label_b, auto_crop)
    na, nb, inter, n = _overlap_counts(eng, a, b)
    true_neg, false_pos = n - (na + nb - inter), nb - inter
    return float((1.0 * true_neg) / (true_neg + false_pos))


def compute_metric_sensitivity(label_a, label_b, auto_crop=True):
    """comparison.py:245-270."""
    eng, a, b = _pair(label_a, label_b, auto_crop)
    na, nb, inter, _ = _overlap_counts(eng, a, b)
    return float((1.0 * inter) / (inter + (na - inter)))


# ---------------------------------------------------------------------------------------------------------------------
# surface metrics
# ---------------------------------------------------------------------------------------------------------------------
def _label_intensity_statistics(eng, contour, distance):
    """itk::LabelIntensityStatisticsImageFilter for label 1 of ``contour`` over |distance|:
    (mean, maximum, standard deviation, median, number of pixels)."""
    vals = np.abs(_values_at(eng, distance, contour)).astype(np.float64)  # sitk.Abs on the values that are used
    if vals.size == 0:
        raise RuntimeError("LabelIntensityStatisticsImageFilter: label 1 is not present in the contour image")  # ITK's GetMean(1) failure
    dmin, dmax = eng.minmax(distance)
    fmax = max(abs(dmin), abs(dmax))        # range of the Abs image: its maximum ...
    fmin = 0.0 if dmin <= 0.0 <= dmax else min(abs(dmin), abs(dmax))  # ... and its minimum (0 on the reference's contour)
    width = (fmax - fmin) / _HISTOGRAM_BINS
    idx = np.minimum(np.floor((vals - fmin) / width).astype(np.int64), _HISTOGRAM_BINS - 1) if width > 0 else np.zeros(vals.size, np.int64)
    cum = np.cumsum(np.bincount(idx, minlength=_HISTOGRAM_BINS))
    median = fmin + (int(np.argmax(cum >= vals.size / 2)) + 0.5) * width
    std = vals.std(ddof=1) if vals.size > 1 else 0.0
    return vals.mean(), vals.max(), std, median, vals.size


def _surface_statistics(eng, a, b):
    rows = []
    for la, lb in ((a, b), (b, a)):
        distance = eng.signed_maurer_distance_map(la, inside_is_positive=False, squared_distance=False, use_image_spacing=True)
        rows.append(_label_intensity_statistics(eng, eng.label_contour(lb, fully_connected=False), distance))
    return rows


def _hausdorff(eng, a, b):
    """itk::HausdorffDistanceImageFilter: max over the voxels of one label of max(signed distance to the other, 0), both ways.
    ITK's directed filter keeps its distance map in double; the Float32 map is therefore taken SQUARED (an exact sum of squares for
    spacings like (1, 1, 2)) and the root of the maximum is formed in double on the host."""
    out = []
    for la, lb in ((a, b), (b, a)):
        d2 = eng.signed_maurer_distance_map(lb, inside_is_positive=False, squared_distance=True, use_image_spacing=True)
        # outside la the masked image is 0, so its maximum is max(0, max over la) -- the filter's max(d, 0)
        out.append(float(np.sqrt(max(0.0, eng.minmax(eng.mask_image(d2, la))[1]))))
    return max(out)


def compute_surface_dsc(label_a, label_b, tau=3.0):
    """Surface Dice (Nikolov et al.; comparison.py:35-72): the share of the two fully-connected contours that lies within ``tau``
    millimetres of the other contour."""
    eng, a, b = _pair(label_a, label_b, False)
    a_contour, b_contour = eng.label_contour(a, fully_connected=True), eng.label_contour(b, fully_connected=True)  # BinaryContour, FullyConnectedOn
    dist_to_a = eng.signed_maurer_distance_map(a_contour, inside_is_positive=False, squared_distance=False, use_image_spacing=True)
    dist_to_b = eng.signed_maurer_distance_map(b_contour, inside_is_positive=False, squared_distance=False, use_image_spacing=True)
    at_b, at_a = _values_at(eng, dist_to_a, b_contour), _values_at(eng, dist_to_b, a_contour)
    return (int((at_b <= tau).sum()) + int((at_a <= tau).sum())) / (at_a.size + at_b.size)


def compute_surface_metrics(label_a, label_b, verbose=False):
    """hausdorffDistance, hausdorffDistance95, mean / median / maximum / sigma surface distance and surfaceDSC
    (comparison.py:75-141)."""
    eng, a, b = _pair(label_a, label_b, False)
    rows = _surface_statistics(eng, a, b)
    mean_sd_list, max_sd_list, std_sd_list, median_sd_list, num_points = (list(c) for c in zip(*rows))
    if verbose:
        print("        Boundary points:  {0}  {1}".format(num_points[0], num_points[1]))
    mean_surf_dist = np.dot(mean_sd_list, num_points) / np.sum(num_points)
    return {
        "hausdorffDistance": _hausdorff(eng, a, b),
        "hausdorffDistance95": np.percentile(max_sd_list, 95),
        "meanSurfaceDistance": mean_surf_dist,
        "medianSurfaceDistance": np.mean(median_sd_list),
        "maximumSurfaceDistance": np.max(max_sd_list),
        # the reference's expression, kept as it is (no division by the number of points, comparison.py:117-127)
        "sigmaSurfaceDistance": np.sqrt(np.dot(num_points, np.add(np.square(std_sd_list), np.square(np.subtract(mean_sd_list, mean_surf_dist))))),
        "surfaceDSC": compute_surface_dsc(a, b),
    }


def compute_metric_masd(label_a, label_b, auto_crop=True):
    """Mean absolute surface distance (comparison.py:273-312)."""
    eng, a, b = _pair(label_a, label_b, auto_crop)
    if _count(eng, a) == 0 or _count(eng, b) == 0:
        return np.nan
    rows = _surface_statistics(eng, a, b)
    means, counts = [r[0] for r in rows], [r[4] for r in rows]
    return float(np.dot(means, counts) / np.sum(counts))


def compute_metric_hd(label_a, label_b, auto_crop=True):
    """Hausdorff distance (comparison.py:315-343)."""
    eng, a, b = _pair(label_a, label_b, auto_crop)
    if _count(eng, a) == 0 or _count(eng, b) == 0:
        return np.nan
    return _hausdorff(eng, a, b)


# ---------------------------------------------------------------------------------------------------------------------
# added path length
# ---------------------------------------------------------------------------------------------------------------------
def compute_apl(label_ref, label_test, distance_threshold_mm=3):
    """Added path length per axial slice, in voxels (comparison.py:346-387): the contour voxels of the reference that are not
    within ``distance_threshold_mm`` (in-plane, rounded up to whole voxels) of the test contour.  All slices are processed at once:
    slice-wise contours, an in-plane disc dilation, MaskNegated; the list holds one count per slice that contains either label."""
    eng, ref, test = _pair(label_ref, label_test, False)
    distance = int(np.ceil(distance_threshold_mm / np.mean(ref.GetSpacing()[:2])))
    ref_contour, test_contour = eng.label_contour_slicewise(ref), eng.label_contour_slicewise(test)
    if distance_threshold_mm > 0:
        test_contour = eng.binary_dilate(test_contour, lu.ball_offsets([distance, distance, 0]))  # sitk.BinaryDilate on the 2-D slice: a disc
    agreement_free = eng.binary_threshold(test_contour, -1.0, 0.5)  # 1 where the (dilated) test contour is 0
    added_path = eng.mask_image(ref_contour, agreement_free)        # sitk.MaskNegated(ref_contour, test_contour)
    with torch.cuda.stream(eng.stream):  # plumbing: per-slice sums of three label volumes
        per_slice = torch.stack([added_path.tensor.sum(dim=(1, 2), dtype=torch.int64), ref.tensor.sum(dim=(1, 2), dtype=torch.int64),
                                 test.tensor.sum(dim=(1, 2), dtype=torch.int64)]).cpu().numpy()
    return [per_slice[0, i] for i in range(per_slice.shape[1]) if per_slice[1, i] + per_slice[2, i] != 0]


def compute_metric_total_apl(label_ref, label_test, distance_threshold_mm=3):
    """Total (slice-wise) added path length in mm (comparison.py:390-409)."""
    added_path_length_list = compute_apl(label_ref, label_test, distance_threshold_mm=distance_threshold_mm)
    return np.sum(added_path_length_list) * np.mean(label_ref.GetSpacing()[:2])


def compute_metric_mean_apl(label_ref, label_test, distance_threshold_mm=3):
    """Mean (slice-wise) added path length in mm (comparison.py:412-431)."""
    added_path_length_list = compute_apl(label_ref, label_test, distance_threshold_mm=distance_threshold_mm)
    return np.mean(added_path_length_list) * np.mean(label_ref.GetSpacing()[:2])
