"""
Iterative atlas removal, drop-in for

    run_iar                           platipy/imaging/label/iar.py:59-301
    evaluate_distance_to_reference    platipy/imaging/label/projection.py:67-92

The per-atlas work -- consensus by ``combine_labels``, ``process_probability_image``, the signed Maurer distance map of the
test label and the contour of the consensus -- runs on the device; what comes back per atlas is the vector of distances at
the consensus surface voxels (raster order), a few thousand numbers.  The statistics on those vectors (modified z-scores,
the Gaussian fit of their histogram, the Q metric and the outlier rule) are host numpy / scipy exactly as in the reference.

``project_on_sphere=True`` is not usable in the reference at this commit (projection.py:42-43 calls ``.mean`` on the tuple
``np.where`` returns, an AttributeError); the same error is raised here.
"""
from __future__ import annotations

import logging

import numpy as np
import torch

from .engine import Engine
from .fusion import combine_labels

logger = logging.getLogger(__name__)


def median_absolute_deviation(data, axis=None):
    """Median absolute deviation (iar.py:36-41)."""
    return np.median(np.abs(data - np.median(data, axis=axis)), axis=axis)


def gaussian_curve(x, a, m, s):
    """a * N(m, s) sampled at x (iar.py:44-56)."""
    from scipy.stats import norm

    return a * norm.pdf(x, loc=m, scale=s)


def _device_mask(eng, image, threshold):
    """process_probability_image(image, threshold) staying on the device (iar.py:124,147)."""
    d = eng.to_device(image)
    if d.np_dtype not in (np.dtype(np.float32), np.dtype(np.float64)):
        d = eng.cast(d, np.float32)  # binary labels: p / max(p) is exact for {0, 1}
    return eng.process_probability(d, threshold)


def _surface_distances(eng, reference_mask, test_mask, resample_factor=1):
    # projection.py:80-90: |SignedMaurerDistanceMap(test)| sampled on LabelContour(reference) == 1, every resample_factor-th
    distance = eng.signed_maurer_distance_map(test_mask, inside_is_positive=False, squared_distance=False, use_image_spacing=True)
    surface = eng.label_contour(reference_mask, fully_connected=False)
    with torch.cuda.stream(eng.stream):
        host = distance.tensor[surface.tensor == 1].cpu()  # plumbing: stream compaction in raster order, like the numpy mask index
    eng.synchronize()
    return np.abs(host.numpy())[::resample_factor]  # sitk.Abs of the few values that are used


def evaluate_distance_to_reference(reference_volume, test_volume, resample_factor=1):
    """Distance from the surface of a test volume to every surface voxel of a reference volume (projection.py:67-92)."""
    eng = Engine.get()
    ref = eng.cast(eng.to_device(reference_volume), np.uint8)
    test = eng.cast(eng.to_device(test_volume), np.uint8)
    return _surface_distances(eng, ref, test, resample_factor)


def _z_scores(own, others, statistic):
    """Modified z-scores of one atlas against the rest (iar.py:171-199)."""
    kind = statistic.lower()
    if kind == "std":
        centre, spread = np.mean(others, axis=0), np.std(others, axis=0)
        if np.any(spread == 0):
            logger.info("    Std Dev zero count: %d", np.sum(spread == 0))
            spread[spread == 0] = spread.mean()
    elif kind == "mad":
        centre, spread = np.median(others, axis=0), 1.4826 * median_absolute_deviation(others, axis=0)
        if np.any(~np.isfinite(spread)):
            logger.info("Error in MAD")
        if np.any(spread == 0):
            logger.info("    MAD zero count: %d", np.sum(spread == 0))
            spread[spread == 0] = np.median(spread)
    else:
        logger.error(" z_score must be one of: MAD, STD")
        raise ValueError("z_score must be one of: MAD, STD")
    with np.errstate(divide="ignore", invalid="ignore"):  # a zero spread gives inf / nan exactly as in the reference expression
        return np.ravel((own - centre) / spread)


def _q_value(z_scores):
    """Excess area of the z-score histogram over its Gaussian fit, weighted by z^2 (iar.py:211-230)."""
    from scipy.optimize import curve_fit

    density, edges = np.histogram(z_scores, bins=np.linspace(-15, 15, 501), density=True)
    centres = (edges[1:] + edges[:-1]) / 2.0
    try:
        popt, _ = curve_fit(f=gaussian_curve, xdata=centres, ydata=density)
        ideal = gaussian_curve(centres, *popt)
    except (RuntimeError, ValueError):
        logger.debug("IAR couldnt fit curve, estimating with sampled statistics.")
        ideal = gaussian_curve(centres, a=1, m=density.mean(), s=density.std())
    trapezoid = getattr(np, "trapezoid", None) or np.trapz
    return np.float64(trapezoid(np.abs(density - ideal) * np.abs(centres) ** 2, centres))


def _outlier_limit(q_values, method, factor, min_best_atlases):
    # iar.py:232-249: the (at most) three worst results are left out of the estimate
    finite = [r for r in q_values if ~np.isnan(r) and np.isfinite(r)]
    best = np.sort(finite)[: max([min_best_atlases, len(finite) - 3])]
    kind = method.lower()
    if kind == "iqr":
        return np.percentile(best, 75, axis=0) + factor * np.subtract(*np.percentile(best, [75, 25], axis=0))
    if kind == "std":
        return np.mean(best, axis=0) + factor * np.std(best, axis=0)
    logger.error(" outlier_method must be one of: IQR, STD")
    raise SystemExit  # the reference calls sys.exit() here (iar.py:249)


def run_iar(atlas_set, reference_structure, smooth_distance_maps=False, smooth_sigma=1, z_score_statistic="MAD", outlier_method="IQR",
            min_best_atlases=10, outlier_factor=1.5, iteration=0, single_step=False, project_on_sphere=False, label="DIR"):
    """Perform iterative atlas removal on the atlas_set (iar.py:59-301); returns the atlas set without the outliers."""
    if iteration == 0:
        logger.info("Iterative atlas removal: ")
        logger.info("  Beginning process")
    if project_on_sphere:
        # projection.py:42-43 in the reference: np.where(...) is a tuple
        raise AttributeError("'tuple' object has no attribute 'mean'")
    eng = Engine.get()
    remaining = list(atlas_set.keys())

    # consensus surface (iar.py:90-92,147)
    probability_label = combine_labels(atlas_set, reference_structure, label=label)[reference_structure]
    if len(remaining) < 12:  # iar.py:105-112 (the "< 7" branch of the reference is unreachable)
        resample_factor = 5
    else:
        resample_factor = 1
    reference_mask = _device_mask(eng, probability_label, 0.95)

    g_vals = []
    logger.info("  Calculating surface distance maps: ")
    for test_id in remaining:
        logger.info("    %s", test_id)
        test_mask = _device_mask(eng, atlas_set[test_id][label][reference_structure], 0.1)
        g_vals.append(_surface_distances(eng, reference_mask, test_mask, resample_factor))

    q_results = {}
    for i, test_id in enumerate(remaining):
        others = g_vals[:i] + g_vals[i + 1:]
        q_results[test_id] = _q_value(_z_scores(g_vals[i], others, z_score_statistic))

    limit = _outlier_limit(list(q_results.values()), outlier_method, outlier_factor, min_best_atlases)
    logger.info("  Analysing results")
    logger.info("   Outlier limit: %6.3f", limit)
    keep = []
    for idx, q in q_results.items():
        accept = q <= limit
        logger.info("      %s: Q = %6.3f [%s]", idx, q, "KEEP" if accept else "REMOVE")
        if accept:
            keep.append(idx)

    if len(keep) < len(remaining):
        logger.info("  Step %d Complete: removed %d", iteration, len(remaining) - len(keep))
        reduced = {i: atlas_set[i] for i in keep}
        if single_step:
            return reduced
        return run_iar(atlas_set=reduced, reference_structure=reference_structure, smooth_distance_maps=smooth_distance_maps,
                       smooth_sigma=smooth_sigma, z_score_statistic=z_score_statistic, outlier_method=outlier_method,
                       min_best_atlases=min_best_atlases, outlier_factor=outlier_factor, iteration=iteration + 1,
                       project_on_sphere=project_on_sphere, label=label)
    logger.info("  End point reached. Keeping:\n   %s", keep)
    return atlas_set


# results of the statistics stage, exposed for tests
q_value = _q_value
z_scores = _z_scores
