"""
Drop-in replacements for the reference's label-fusion entry points (platipy/imaging/label/fusion.py):

    compute_weight_map      fusion.py:56-202   (vote types: unweighted, global, local, block; normalise)
    combine_labels          fusion.py:239-292  (weighted vote -> DiscreteGaussian -> RescaleIntensity -> Threshold)
    combine_labels_staple   fusion.py:205-236  (BinaryThreshold -> STAPLE -> RescaleIntensity -> Threshold)
    process_probability_image fusion.py:295-328 (normalise -> BinaryThreshold -> BinaryFillhole -> largest component)

plus the sharded forms used by ``platipy_b200.multiatlas``: every rank accumulates the votes of its local
atlases and ONE all-reduce (NCCL on GPUs, gloo in the CPU tests of the host logic) exchanges the per-voxel
vote volume before the replicated finalisation.
"""
from __future__ import annotations

import numpy as np

from . import sitk_compat as sk
from .engine import DeviceImage, Engine

VOTE_TYPES = {"unweighted": 0, "global": 1, "local": 2}

DEFAULT_VOTE_PARAMS = {
    "sigma": 2.0,
    "epsilon": 1e-5,
    "factor": 1e12,
    "gain": 6,
    "blockSize": 5,
    "normalise": False,
    "patch_window_mm": 25,
    "resampled_voxel_size_mm": 3,
    "correlation_function": lambda x: x + 1,
}


def mutual_information(arr_a, arr_b, bins=64):
    """Histogram-based mutual information of two flattened numpy arrays (fusion.py:26-53).  A host utility on host arrays in
    the reference as well (no image, no ITK filter behind it): the same numpy expression."""
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        p_ab, _, _ = np.histogram2d(arr_a, arr_b, bins=bins, density=True)
        log_p = np.log(p_ab / np.outer(p_ab.sum(axis=0), p_ab.sum(axis=1)))
    log_p[~np.isfinite(log_p)] = 0
    return (p_ab * log_p).sum()


def _back(eng, dimg, like):
    if isinstance(like, DeviceImage):
        eng.release_to_caller()
        return dimg
    return sk.from_native(eng.to_host(dimg), like)


def compute_weight_map(target_image, moving_image, vote_type="unweighted", vote_params=DEFAULT_VOTE_PARAMS):
    """Weight map of one atlas (fusion.py:56-202): ``unweighted``, ``global``, ``local`` and ``block`` votes, with the
    ``normalise`` option (bool or mask image) of the last two.  ``vote_params=None`` is valid for ``unweighted`` only,
    as in the reference (multiatlas/run.py:92-96)."""
    eng = Engine.get()
    vt = vote_type.lower()
    if vt not in VOTE_TYPES and vt not in ("block", "patch_correlation"):
        # the reference falls through its if / elif chain and fails on the unbound weight_map (fusion.py:151-202)
        raise UnboundLocalError(f"vote_type {vote_type!r}: local variable 'weight_map' referenced before assignment")
    t, m = eng.to_device(target_image), eng.to_device(moving_image)
    # fusion.py:76-80: cast to Float32 unless the pixel id is 6
    t, m = eng.cast(t, np.float32), eng.cast(m, np.float32)
    if t.GetSize() != m.GetSize():
        raise RuntimeError("compute_weight_map: target and moving images do not occupy the same grid")
    normalise = False
    if vt == "patch_correlation":
        w = _patch_correlation(eng, t, m, vote_params, target_image)
    elif vt == "block":
        factor, gain, block_size = float(vote_params["factor"]), float(vote_params["gain"]), vote_params["blockSize"]
        normalise = vote_params["normalise"]
        if isinstance(block_size, int):
            block_size = (block_size,) * 3  # fusion.py:185-186
        w = eng.weight_map_block(t, m, block_size, factor, gain)
    else:
        factor = sigma = eps = 0.0
        if vt == "global":
            factor = float(vote_params["factor"])
        elif vt == "local":
            sigma, eps = float(vote_params["sigma"]), float(vote_params["epsilon"])
            normalise = vote_params["normalise"]
        w = eng.weight_map(t, m, VOTE_TYPES[vt], factor=factor, sigma=sigma, epsilon=eps)
    # fusion.py:171-177,196-200
    if isinstance(normalise, bool):
        if normalise:
            eng.normalise_by_max(w)
    else:
        mask = eng.cast(eng.to_device(normalise), np.uint8)
        eng.normalise_by_max(w, mask)
    return _back(eng, w, target_image)


def _patch_correlation(eng, t, m, vote_params, like):
    """fusion.py:82-146: both images resampled to isotropic voxels (linear, no smoothing), Pearson correlation of the two over
    a cubic window around every voxel of that grid (one kernel instead of the reference's Python loop over patches),
    resampled back onto the target grid (linear, default 0), ``correlation_function`` applied, cast to Float32."""
    from .registration import smooth_and_resample

    voxel_size = vote_params["resampled_voxel_size_mm"]
    t_res = smooth_and_resample(t, isotropic_voxel_size_mm=voxel_size)
    m_res = smooth_and_resample(m, isotropic_voxel_size_mm=voxel_size)
    eng.wait_caller()
    window_zyx = [int(vote_params["patch_window_mm"] / i) for i in t_res.GetSpacing()[::-1]]  # fusion.py:97
    corr = eng.patch_correlation(t_res, m_res, window_zyx[::-1])
    corr = eng.resample(corr, t, None, sk.sitkLinear, 0.0)  # sitk.Resample(corr_img, target_image)
    fn = vote_params["correlation_function"]
    try:
        w = fn(corr)  # device arithmetic: x + 1, abs(x), 2 * x ...
    except (TypeError, AttributeError):
        # a function written against the SimpleITK API (sitk.Abs(x) ...): hand it the image in the caller's representation
        w = fn(sk.from_native(eng.to_host(corr), like))
    return eng.cast(eng.to_device(w), np.float32)


def _structure_names(structure_name):
    if isinstance(structure_name, str):
        return [structure_name]
    return list(structure_name)


def accumulate_votes(eng, atlas_set, s_name, label="DIR", num=None, den=None):
    """Local part of combine_labels for one structure: num += w * L, den += w over the atlases of
    ``atlas_set`` that contain the structure (fusion.py:255-276), float32 arithmetic in atlas order."""
    first = num is None
    ref = None
    for case_id in atlas_set:
        entry = atlas_set[case_id][label]
        if s_name not in entry:
            continue
        lab = eng.cast(eng.to_device(entry[s_name]), np.uint8)
        # fusion.py:263-272 multiplies the Float32 weight map with the label cast to Float32: both on one grid (ITK raises
        # "Inputs do not occupy the same physical space" otherwise); a weight map of another pixel type is cast like sitk would
        w = eng.cast(eng.to_device(entry["Weight Map"]), np.float32)
        if w.GetSize() != lab.GetSize():
            raise RuntimeError(f"combine_labels: weight map {w.GetSize()} and label {s_name!r} {lab.GetSize()} of atlas {case_id!r} "
                               "do not occupy the same grid")
        if num is not None and tuple(num.shape) != tuple(lab.tensor.shape):
            raise RuntimeError(f"combine_labels: label {s_name!r} of atlas {case_id!r} has size {lab.GetSize()}, the votes so far "
                               f"{tuple(num.shape)[::-1]}")
        if ref is None:
            ref = lab
        if first:
            num = eng.empty(lab.tensor.shape, np.float32)
            den = eng.empty(lab.tensor.shape, np.float32)
        eng.vote_accumulate(lab, w, num, den, first)
        first = False
    return num, den, ref


def combine_labels(atlas_set, structure_name, label="DIR", threshold=1e-4, smooth_sigma=1.0, process_group=None, reference_image=None):
    """Combine labels using weight maps (fusion.py:239-292).  With ``process_group`` (torch.distributed; ``True`` = the default
    group) every rank passes its LOCAL atlases and the vote volumes are summed with one all-reduce per structure;
    ``reference_image`` (any image on the common grid, e.g. the target) is then required on ranks that may hold no atlas with a
    structure -- it gives the grid of their all-zero contribution.  ``multiatlas.run_segmentation`` uses the structure-sharded
    reduce-scatter form of this exchange instead."""
    eng = Engine.get()
    out = {}
    any_like = None
    for s_name in _structure_names(structure_name):
        num, den, ref = accumulate_votes(eng, atlas_set, s_name, label)
        if process_group is not None:
            num, den, ref = _allreduce_votes(eng, num, den, ref, reference_image, process_group)
        if ref is None:
            raise KeyError(s_name)
        for case_id in atlas_set:
            if s_name in atlas_set[case_id][label]:
                any_like = atlas_set[case_id][label][s_name]
                break
        if any_like is None:
            any_like = reference_image
        prob = eng.vote_finalize(num, den, ref, smooth_sigma * smooth_sigma, threshold)
        out[s_name] = _back(eng, prob, any_like if any_like is not None else ref)
    return out


def _allreduce_votes(eng, num, den, ref, reference_image, process_group):
    import torch
    import torch.distributed as dist

    if num is None:
        # this rank holds no atlas with the structure: it still has to enter the collective, with zeros on the common grid
        if reference_image is None:
            raise ValueError("combine_labels(process_group=...): this rank holds no atlas with the structure; pass reference_image "
                             "(an image on the common grid, e.g. the target) so that it can contribute zeros to the all-reduce")
        ref = eng.to_device(reference_image)
        num = eng.zeros(ref.tensor.shape, np.float32)
        den = eng.zeros(ref.tensor.shape, np.float32)
    group = process_group if process_group is not True else None
    with torch.cuda.stream(eng.stream):
        dist.all_reduce(num, group=group)
        dist.all_reduce(den, group=group)
    return num, den, ref


def combine_labels_staple(label_list_dict, threshold=1e-4):
    """Combine labels using STAPLE (fusion.py:205-236)."""
    eng = Engine.get()
    names = np.unique([n for d in label_list_dict.values() for n in d.keys()])
    out = {}
    for s_name in names:
        imgs = [label_list_dict[i][s_name] for i in label_list_dict]  # KeyError if an atlas lacks it, as in the reference
        dec = []
        for im in imgs:
            d = eng.to_device(im)
            # sitk.BinaryThreshold(lowerThreshold=0.5) (upper 255) -> UInt8 {0, 1}
            dec.append(eng.binary_threshold(d, 0.5, 255.0))
        w, _info = eng.staple(dec, threshold=threshold, rescale=True)
        out[str(s_name)] = _back(eng, w, imgs[0])
    return out


def process_probability_image(probability_image, threshold=0.5):
    """Generate a mask given a probability image, performing some basic post processing as well
    (fusion.py:295-328): p / max(p) -> BinaryThreshold(lowerThreshold=threshold) -> BinaryFillhole ->
    ConnectedComponent -> largest object -> UInt8.  One device-resident call; integer work, bit-exact.
    A numpy array is accepted like in the reference (fusion.py:301-302: identity geometry)."""
    eng = Engine.get()
    if isinstance(probability_image, np.ndarray):
        probability_image = sk.Image(probability_image)
    d = eng.to_device(probability_image)
    if d.np_dtype not in (np.dtype(np.float32), np.dtype(np.float64)):
        # integer probability images: the reference's image / float division keeps the integer pixel type; the fused
        # probabilities this path produces are always Float32 (combine_labels) or Float64 (combine_labels_staple)
        raise NotImplementedError("process_probability_image expects a Float32 or Float64 probability image")
    mask = eng.process_probability(d, threshold)
    return _back(eng, mask, probability_image)
