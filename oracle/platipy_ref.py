"""
CPU restatement of the reference's Python glue on the hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

PARITY UNPINNED (see ``itk_oracle.c``).  Function names and argument meaning follow the reference
(paths relative to /root/reference):

  smooth_and_resample                          platipy/imaging/registration/utils.py:195-267
  apply_transform                              platipy/imaging/registration/utils.py:148-192
  multiscale_demons                            platipy/imaging/registration/deformable.py:31-187
  fast_symmetric_forces_demons_registration    platipy/imaging/registration/deformable.py:190-306
  compute_weight_map / combine_labels / combine_labels_staple   platipy/imaging/label/fusion.py:56-292

Every ITK filter call in those functions is replaced by the corresponding routine of ``itk_oracle``.
Images are ``platipy_b200.sitk_compat.Image`` containers (data holders only; no product compute).
"""
from __future__ import annotations

import numpy as np

from platipy_b200 import sitk_compat as sk
from platipy_b200.sitk_compat import Image

from . import itk_oracle as orc


def _like(arr, ref, is_vector=False):
    return Image(arr, ref.GetSpacing(), ref.GetOrigin(), ref.GetDirection(), is_vector)


def _chain_of(transform):
    """sitk transform object -> oracle chain in application order."""
    chain = []
    if transform is None:
        return chain
    for t in transform.flatten():
        if isinstance(t, sk.DisplacementFieldTransform):
            f = t.GetDisplacementField()
            chain.append(("dvf", f.array, orc.geom_of(f)))
        else:
            chain.append(("affine", t.matrix, t.offset))
    return chain


def resample(image, reference_geom_image=None, transform=None, interpolator=sk.sitkLinear, default_value=0.0,
             size=None, spacing=None, origin=None, direction=None):
    """sitk.Resample in the three call forms the reference uses."""
    if reference_geom_image is None:
        reference_geom_image = image
    if size is None:
        gout = orc.geom_of(reference_geom_image)
        sp, og, dr = reference_geom_image.GetSpacing(), reference_geom_image.GetOrigin(), reference_geom_image.GetDirection()
    else:
        gout = orc.make_geom(size, spacing, origin, direction)
        sp, og, dr = spacing, origin, direction
    gin = orc.geom_of(image)
    chain = _chain_of(transform)
    if image.is_vector:
        out = orc.resample_vec3(image.array, gin, gout, chain, default_value)
        return Image(out, sp, og, dr, True)
    if interpolator not in (sk.sitkNearestNeighbor, sk.sitkLinear, sk.sitkBSpline):
        raise NotImplementedError("oracle: nearest-neighbour, linear and B-spline (order 3) interpolation")
    out = orc.resample_scalar(image.array, gin, gout, chain, interpolator, default_value)
    return Image(out, sp, og, dr, False)


def smooth_and_resample(image, isotropic_voxel_size_mm=None, shrink_factor=None, smoothing_sigma=None,
                        interpolator=sk.sitkLinear):
    # utils.py:216-226: variance = sigma^2 (mm^2); max kernel width = int(max(8 * var_i * spacing_i))
    if smoothing_sigma:
        if hasattr(smoothing_sigma, "__iter__"):
            variance = [s * s for s in smoothing_sigma]
        else:
            variance = (smoothing_sigma ** 2,) * 3
        max_width = int(max(8 * v * sp for sp, v in zip(image.GetSpacing(), variance)))
        if image.array.dtype != np.float32:
            raise NotImplementedError("oracle: DiscreteGaussian restated for Float32 images")
        image = _like(orc.discrete_gaussian_f32(image.array, orc.geom_of(image), variance, max_width), image)

    size_o, spacing_o = image.GetSize(), image.GetSpacing()
    # utils.py:231-250
    if shrink_factor and isotropic_voxel_size_mm:
        raise AttributeError("Function must be called with either isotropic_voxel_size_mm or shrink_factor, not both.")
    elif isotropic_voxel_size_mm:
        scale = isotropic_voxel_size_mm * np.ones_like(size_o) / np.array(spacing_o)
        size_n = [int(sz / float(sf) + 0.5) for sz, sf in zip(size_o, scale)]
    elif shrink_factor:
        if isinstance(shrink_factor, list):
            size_n = [int(sz / float(sf) + 0.5) for sz, sf in zip(size_o, shrink_factor)]
        else:
            size_n = [int(sz / float(shrink_factor) + 0.5) for sz in size_o]
    else:
        return image
    # utils.py:252-255: align-corners spacing
    spacing_n = [((so - 1) * sp) / (sn - 1) for so, sp, sn in zip(size_o, spacing_o, size_n)]
    # utils.py:257-267: identity transform, default pixel 0.0, same pixel type
    return resample(image, transform=None, interpolator=interpolator, default_value=0.0, size=size_n,
                    spacing=spacing_n, origin=image.GetOrigin(), direction=image.GetDirection())


def apply_transform(input_image, reference_image=None, transform=None, default_value=0, interpolator=sk.sitkNearestNeighbor):
    # utils.py:174-190; output pixel type = input pixel type (Resample), Cast back is then a no-op
    ref = reference_image if reference_image else input_image
    return resample(input_image, ref, transform, interpolator, default_value)


class DemonsFilter:
    """The method set of sitk.FastSymmetricForcesDemonsRegistrationFilter that the reference uses
    (deformable.py:244-257,143-149,157; utils.py:41), on the oracle's Demons loop."""

    def __init__(self):
        self.std_dev = [1.0, 1.0, 1.0]
        self.iterations = 10
        self.smooth_update = False
        self.smooth_field = True
        self.stats = None

    def SetNumberOfThreads(self, n):
        pass

    def SetSmoothUpdateField(self, flag):
        self.smooth_update = bool(flag)

    def SetSmoothDisplacementField(self, flag):
        self.smooth_field = bool(flag)

    def SetStandardDeviations(self, sd):
        self.std_dev = [float(sd)] * 3 if np.isscalar(sd) else [float(s) for s in sd]

    def GetStandardDeviations(self):
        return tuple(self.std_dev)

    def SetNumberOfIterations(self, n):
        self.iterations = int(n)

    def GetElapsedIterations(self):
        return self.stats["elapsed_iterations"]

    def GetMetric(self):
        return self.stats["metric"]

    def GetRMSChange(self):
        return self.stats["rms_change"]

    def Execute(self, fixed, moving):
        params = orc.demons_params(self.std_dev, self.iterations, smooth_displacement_field=self.smooth_field,
                                   smooth_update_field=self.smooth_update)
        D, self.stats = orc.demons_execute(fixed.array, orc.geom_of(fixed), moving.array, orc.geom_of(moving), params)
        return _like(D, fixed, True)


def multiscale_demons(registration_algorithm, fixed_image, moving_image, initial_transform=None,
                      initial_displacement_field=None, isotropic_resample=None, resolution_staging=None,
                      smoothing_sigmas=None, iteration_staging=None, interp_order=sk.sitkLinear, level_stats=None):
    # deformable.py:67-94: pyramid from the original images for every level
    fixed_images, moving_images = [], []
    for resolution, sigma in zip(resolution_staging, smoothing_sigmas):
        kw = {"isotropic_voxel_size_mm": resolution} if isotropic_resample else {"shrink_factor": resolution}
        fixed_images.append(smooth_and_resample(fixed_image, smoothing_sigma=sigma, interpolator=interp_order, **kw))
        moving_images.append(smooth_and_resample(moving_image, smoothing_sigma=sigma, interpolator=interp_order, **kw))

    # deformable.py:99-125
    if not initial_displacement_field:
        if initial_transform:
            # deformable.py:101-108 sitk.TransformToDisplacementField on the fixed grid
            arr = orc.transform_to_dvf(orc.geom_of(fixed_image), _chain_of(initial_transform))
            initial_displacement_field = _like(arr, fixed_image, True)
        else:
            zeros = np.zeros(fixed_image.array.shape + (3,), dtype=np.float64)
            initial_displacement_field = _like(zeros, fixed_image, True)
    else:
        initial_displacement_field = resample(initial_displacement_field, fixed_image)

    dvf_total = resample(initial_displacement_field, fixed_image)  # :130
    for f_image, m_image, iters in zip(fixed_images, moving_images, iteration_staging):
        dvf_total = resample(dvf_total, f_image)  # :137
        tfm_total = sk.DisplacementFieldTransform(sk.Cast(dvf_total, sk.sitkVectorFloat64))  # :139
        m_image = resample(m_image, m_image, tfm_total, interp_order, 0.0)  # :140
        registration_algorithm.SetNumberOfIterations(iters)  # :143-144
        dvf_iter = registration_algorithm.Execute(f_image, m_image)  # :149
        if level_stats is not None:
            level_stats.append(dict(registration_algorithm.stats, voxels=f_image.GetNumberOfPixels()))
        warped_iter = resample(dvf_iter, dvf_iter, tfm_total)  # :154  Resample(dvf_iter, tfm_total)
        dvf_total = _like(dvf_total.array + warped_iter.array, dvf_total, True)
        sigma = registration_algorithm.GetStandardDeviations()  # :157
        dvf_total = _like(orc.recursive_gaussian_vec3(dvf_total.array, orc.geom_of(dvf_total), sigma), dvf_total, True)  # :158
    return resample(dvf_total, fixed_image)  # :185


def fast_symmetric_forces_demons_registration(fixed_image, moving_image, resolution_staging=[8, 4, 1],
                                              iteration_staging=[10, 10, 10], isotropic_resample=False,
                                              initial_displacement_field=None, regularisation_kernel_mm=1.5,
                                              smoothing_sigma_factor=1, smoothing_sigmas=False, default_value=None,
                                              ncores=1, interp_order=sk.sitkLinear, verbose=False, level_stats=None):
    moving_type = moving_image.GetPixelID()
    # deformable.py:238-241 (pixel id 6 is Int64: the quirk is kept)
    if fixed_image.GetPixelID() != 6:
        fixed_image = sk.Cast(fixed_image, sk.sitkFloat32)
    if moving_image.GetPixelID() != 6:
        moving_image = sk.Cast(moving_image, sk.sitkFloat32)
    reg = DemonsFilter()
    reg.SetNumberOfThreads(ncores)
    reg.SetSmoothUpdateField(True)
    reg.SetSmoothDisplacementField(True)
    # deformable.py:253-257: voxel-unit sigmas from the FULL-resolution spacing
    reg.SetStandardDeviations((np.array(regularisation_kernel_mm) / np.array(fixed_image.GetSpacing())).tolist())
    if not smoothing_sigmas:
        smoothing_sigmas = [i * smoothing_sigma_factor for i in resolution_staging]
    dvf = multiscale_demons(reg, fixed_image, moving_image, resolution_staging=resolution_staging,
                            smoothing_sigmas=smoothing_sigmas, iteration_staging=iteration_staging,
                            isotropic_resample=isotropic_resample, initial_displacement_field=initial_displacement_field,
                            interp_order=interp_order, level_stats=level_stats)
    # deformable.py:286-293
    if default_value is None:
        default_value = 0
        if moving_image.array.min() <= -1000:
            default_value = -1000
    tfm = sk.DisplacementFieldTransform(sk.Cast(dvf, sk.sitkVectorFloat64))
    registered = resample(moving_image, fixed_image, tfm, interp_order, default_value)  # :281-301
    registered = sk.Cast(registered, moving_type)  # :303-304
    return registered, tfm, dvf


def box_mean_f32(arr, radius_xyz):
    """sitk.BoxMean on a Float32 [z, y, x] array: mean over the window [i - r, i + r] cropped to the image
    (itk::BoxMeanImageFilter divides the box sum by the number of pixels inside), sums in double, Float32 result."""
    a = arr.astype(np.float64)
    for axis, r in zip((2, 1, 0), radius_xyz):
        n = a.shape[axis]
        c = np.concatenate([np.zeros_like(np.take(a, [0], axis=axis)), np.cumsum(a, axis=axis)], axis=axis)
        idx = np.arange(n)
        hi, lo = np.minimum(idx + r, n - 1) + 1, np.maximum(idx - r, 0)
        a = np.take(c, hi, axis=axis) - np.take(c, lo, axis=axis)
    cnt = np.ones((), np.float64)
    for axis, r in zip((2, 1, 0), radius_xyz):
        n = arr.shape[axis]
        idx = np.arange(n)
        ln = (np.minimum(idx + r, n - 1) - np.maximum(idx - r, 0) + 1).astype(np.float64)
        shape = [1, 1, 1]
        shape[axis] = n
        cnt = cnt * ln.reshape(shape)
    return (a / cnt).astype(np.float32)


def _normalise(w, normalise):
    # fusion.py:171-177: bool -> divide by the global maximum; image -> by the maximum inside the mask (sitk.Mask zeroes outside)
    if isinstance(normalise, bool):
        if normalise:
            w = (w.astype(np.float64) / float(w.max())).astype(np.float32)
        return w
    mask = normalise.array != 0
    mx = float(np.where(mask, w, np.float32(0)).max())
    return (w.astype(np.float64) / mx).astype(np.float32)


def compute_weight_map(target_image, moving_image, vote_type="unweighted", vote_params=None):
    # fusion.py:76-80,148-202 (unweighted / global / local / block)
    if target_image.GetPixelID() != 6:
        target_image = sk.Cast(target_image, sk.sitkFloat32)
    if moving_image.GetPixelID() != 6:
        moving_image = sk.Cast(moving_image, sk.sitkFloat32)
    t, m = target_image.array, moving_image.array
    sq = ((t.astype(np.float64) - m.astype(np.float64)) ** 2).astype(np.float32)
    vt = vote_type.lower()
    if vt == "unweighted":
        w = (t * np.float32(0.0) + np.float32(1.0)).astype(np.float32)
    elif vt == "global":
        gw = vote_params["factor"] / sq.sum(dtype=np.float64)
        w = (t * np.float32(0.0) + np.float32(gw)).astype(np.float32)
    elif vt == "local":
        sigma, eps = vote_params["sigma"], vote_params["epsilon"]
        raw = orc.discrete_gaussian_f32(sq, orc.geom_of(target_image), sigma * sigma)
        # image + double constant: sum in double, cast to Float32; sitk.Pow: std::pow in double, cast to Float32
        w = np.power((raw.astype(np.float64) + eps).astype(np.float32).astype(np.float64), -1.0).astype(np.float32)
        w = _normalise(w, vote_params.get("normalise", False))
    elif vt == "block":
        # fusion.py:179-200
        factor, gain, block_size = vote_params["factor"], vote_params["gain"], vote_params["blockSize"]
        if isinstance(block_size, int):
            block_size = (block_size,) * 3
        raw = box_mean_f32(sq, block_size)
        with np.errstate(divide="ignore", over="ignore", invalid="ignore"):
            inv = np.power(raw.astype(np.float64), -1.0).astype(np.float32)
            pw = np.power(inv.astype(np.float64), abs(gain / 2.0)).astype(np.float32)
            w = (pw.astype(np.float64) * factor).astype(np.float32)
        w = _normalise(w, vote_params.get("normalise", False))
    elif vt == "patch_correlation":
        w = _patch_correlation(target_image, moving_image, vote_params)
    else:
        raise NotImplementedError(vote_type)
    return _like(w, target_image)


def _patch_correlation(target_image, moving_image, vote_params):
    """fusion.py:82-146 with numpy's sliding_window_view in place of skimage's view_as_windows and scipy.stats.pearsonr on
    Float64 copies of the patches (the arithmetic pearsonr uses for Float32 data under the reference's pinned numpy 1.24.4 /
    scipy 1.9.3: dtype = type(1.0 + x[0] + y[0]) is float64 there)."""
    import warnings

    from numpy.lib.stride_tricks import sliding_window_view
    from scipy.stats import pearsonr

    voxel_size = vote_params["resampled_voxel_size_mm"]
    t_res = smooth_and_resample(target_image, isotropic_voxel_size_mm=voxel_size)
    m_res = smooth_and_resample(moving_image, isotropic_voxel_size_mm=voxel_size)
    arr_t, arr_m = t_res.array, m_res.array
    arr_mask = 0 * arr_t + 1
    window = [int(vote_params["patch_window_mm"] / i) for i in t_res.GetSpacing()[::-1]]
    padder = [((i - 1) // 2, i // 2) for i in window]
    arr_t, arr_m, arr_mask = np.pad(arr_t, padder), np.pad(arr_m, padder), np.pad(arr_mask, padder)
    v_t, v_m, v_k = (sliding_window_view(a, window) for a in (arr_t, arr_m, arr_mask))
    shape = v_t.shape[:3]
    corr = np.empty(int(np.prod(shape)), dtype=np.float64)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for i, idx in enumerate(np.ndindex(*shape)):
            keep = v_k[idx].ravel() != 0
            corr[i] = pearsonr(v_t[idx].ravel()[keep].astype(np.float64), v_m[idx].ravel()[keep].astype(np.float64))[0]
    corr = corr.reshape(t_res.GetSize()[::-1])
    corr[np.isnan(corr)] = 0
    corr_img = resample(_like(corr, t_res), target_image)  # sitk.Resample(corr_img, target_image): linear, default 0
    w = vote_params["correlation_function"](corr_img)
    return w.array.astype(np.float32)


def combine_labels(atlas_set, structure_name, label="DIR", threshold=1e-4, smooth_sigma=1.0):
    # fusion.py:239-292
    names = [structure_name] if isinstance(structure_name, str) else list(structure_name)
    out = {}
    for s in names:
        cases = [c for c in atlas_set if s in atlas_set[c][label]]
        ws = [atlas_set[c][label]["Weight Map"].array for c in cases]
        ls = [atlas_set[c][label][s].array for c in cases]
        ref = atlas_set[cases[0]][label][s]
        out[s] = _like(orc.combine_labels_f32(ls, ws, orc.geom_of(ref), smooth_sigma * smooth_sigma, threshold), ref)
    return out


def _binary_threshold(arr, lower, upper):
    """BinaryThresholdImageFilter -> UInt8 {0, 1}; semantic switch binary_threshold_in_pixel_type: the bounds are cast to an
    integer pixel type (clamped to its range, truncated) before the comparison instead of being compared as real numbers."""
    if orc.get_semantic("binary_threshold_in_pixel_type") and arr.dtype.kind in "iu":
        info = np.iinfo(arr.dtype)
        lower, upper = (int(min(max(v, info.min), info.max)) for v in (lower, upper))
    return ((arr >= lower) & (arr <= upper)).astype(np.uint8)


def combine_labels_staple(label_list_dict, threshold=1e-4):
    # fusion.py:205-236
    names = np.unique([n for d in label_list_dict.values() for n in d.keys()])
    out = {}
    for s in names:
        imgs = [label_list_dict[i][s] for i in label_list_dict]
        binary = [_binary_threshold(im.array, 0.5, 255) for im in imgs]  # BinaryThreshold(lower=0.5, upper=255)
        W, _, _, _ = orc.staple(binary)
        W = orc.rescale_threshold_f64(W, threshold)
        out[str(s)] = _like(W, imgs[0])
    return out


def process_probability_image(probability_image, threshold=0.5):
    """fusion.py:295-328: normalise by the maximum, BinaryThreshold(lowerThreshold=threshold), BinaryFillhole,
    ConnectedComponent, keep the largest object; returns a UInt8 image."""
    if isinstance(probability_image, np.ndarray):
        probability_image = sk.Image(probability_image)  # fusion.py:301-302
    arr = probability_image.array
    mx = arr.max()
    # sitk image / float: itk::Functor::Div<T, double, T> -- A / B in double, cast back to the pixel type (fusion.py:305)
    if mx != 0:
        norm = (arr.astype(np.float64) / float(mx)).astype(arr.dtype)
    else:
        norm = np.full(arr.shape, np.finfo(arr.dtype).max, dtype=arr.dtype)
    # sitk.BinaryThreshold(lowerThreshold=threshold): upper 255, inside 1, outside 0, UInt8 (fusion.py:308)
    nd = norm.astype(np.float64)
    binary = ((nd >= threshold) & (nd <= 255.0)).astype(np.uint8)
    binary = orc.binary_fillhole(binary)                       # fusion.py:311
    largest, ncomp, _ = orc.largest_component(binary)          # fusion.py:314-326
    if ncomp == 0:
        return _like(binary, probability_image)                # fusion.py:322-323
    return _like(largest, probability_image)


def image_moments(image):
    """itk::ImageMomentsCalculator (behind sitk.CenteredTransformInitializer MOMENTS, linear.py:40-43):
    [sum v, sum v x, sum v y, sum v z] with (x, y, z) the physical position of every voxel."""
    v = image.array.astype(np.float64)
    nz, ny, nx = v.shape
    k, j, i = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    d = np.asarray(image.GetDirection(), np.float64).reshape(3, 3)
    idx = np.stack([i, j, k], axis=-1).astype(np.float64)
    pts = (idx * np.asarray(image.GetSpacing())) @ d.T + np.asarray(image.GetOrigin())
    return np.concatenate([[v.sum()], (pts * v[..., None]).sum(axis=(0, 1, 2))])


def linreg_meansq(fixed, moving, total_matrix, total_offset, initial_matrix, center, fixed_mask=None, moving_mask=None, stride=1):
    return _linreg_sums(fixed, moving, total_matrix, total_offset, initial_matrix, center, fixed_mask, moving_mask, stride, "mean_squares")


def linreg_correlation(fixed, moving, total_matrix, total_offset, initial_matrix, center, fixed_mask=None, moving_mask=None, stride=1):
    """Sums behind itk::CorrelationImageToImageMetricv4 (linear.py:141-146 ``SetMetricAsCorrelation``) over the same samples as
    ``linreg_meansq``: [N, sum F, sum M, sum F^2, sum M^2, sum F M] followed, for each weight w in (1, F, M), by
    s_w = sum w h (3) and S_w = sum w h (x - center)^T (9, row-major), h = initial_matrix^T grad M."""
    return _linreg_sums(fixed, moving, total_matrix, total_offset, initial_matrix, center, fixed_mask, moving_mask, stride, "correlation")


def _bspline3(u):
    a = np.abs(u)
    return np.where(a < 1, (4 - 6 * a * a + 3 * a ** 3) / 6, np.where(a < 2, (2 - a) ** 3 / 6, 0.0))


def _bspline3_derivative(u):
    a = np.abs(u)
    d = np.where(a < 1, -2 * a + 1.5 * a * a, np.where(a < 2, -0.5 * (2 - a) ** 2, 0.0))
    return np.sign(u) * d


def linreg_mattes(fixed, moving, total_matrix, total_offset, initial_matrix, center, fixed_bins, moving_bins, n_bins=50, table=None,
                  fixed_mask=None, moving_mask=None, stride=1):
    """itk::MattesMutualInformationImageToImageMetricv4 (linear.py:145-146) over the samples of ``linreg_meansq``: the joint Parzen
    histogram [fixed bin, moving bin] (zero-order window on the fixed value, cubic B-spline on the moving value, bin indices clamped
    to [2, n_bins - 3]) and the sample count; with ``table`` = log(p / p_M) also the derivative sums [s (3), S (9)] for the weight
    w = sum_m table[f][m] B3'(m - u) (-1 / moving bin size).  *_bins = (bin size, normalised minimum)."""
    fv, mval, h, xc = _linreg_sums(fixed, moving, total_matrix, total_offset, initial_matrix, center, fixed_mask, moving_mask, stride, "raw")
    lo, hi = 2, n_bins - 3
    fi = np.clip(np.floor(fv / fixed_bins[0] - fixed_bins[1]).astype(np.int64), lo, hi)
    term = mval / moving_bins[0] - moving_bins[1]
    mi = np.clip(np.floor(term).astype(np.int64), lo, hi)
    hist = np.zeros((n_bins, n_bins))
    w = np.zeros_like(term)
    for off in (-1, 0, 1, 2):
        b = mi + off
        np.add.at(hist, (fi, b), _bspline3(b - term))
        if table is not None:
            w += np.asarray(table)[fi, b] * _bspline3_derivative(b - term)
    if table is None:
        return hist, float(fv.size)
    w = -w / moving_bins[0]
    wh = w[:, None] * h
    return hist, float(fv.size), np.concatenate([wh.sum(axis=0), (wh[:, :, None] * xc[:, None, :]).sum(axis=0).reshape(9)])


def _linreg_sums(fixed, moving, total_matrix, total_offset, initial_matrix, center, fixed_mask, moving_mask, stride, kind):
    """Mean-squares metric sums of linear_registration (linear.py:141-163: SetMetricAsMeanSquares, linear interpolator,
    REGULAR sampling, optional masks), restated in numpy for every ``stride``-th fixed voxel in raster order:
    [sum (M - F)^2, count, s (3), S (9 row-major)] with w = 2 (M - F) initial_matrix^T grad M, s = sum w,
    S = sum w (x - center)^T.  Points mapping outside the moving buffer ([-0.5, size - 0.5), as
    ImageFunction::IsInsideBuffer) or outside a mask are skipped."""
    F, M = fixed.array.astype(np.float64), moving.array.astype(np.float64)
    nz, ny, nx = F.shape
    q = np.arange(0, F.size, int(stride))
    k, j, i = np.unravel_index(q, F.shape)
    keep = np.ones(q.shape, bool)
    if fixed_mask is not None:
        keep &= fixed_mask.array.reshape(-1)[q] != 0
    df = np.asarray(fixed.GetDirection(), np.float64).reshape(3, 3)
    idx = np.stack([i, j, k], axis=1).astype(np.float64)
    x = (idx * np.asarray(fixed.GetSpacing())) @ df.T + np.asarray(fixed.GetOrigin())
    y = x @ np.asarray(total_matrix, np.float64).reshape(3, 3).T + np.asarray(total_offset, np.float64)
    dm = np.asarray(moving.GetDirection(), np.float64).reshape(3, 3)
    i2p = dm * np.asarray(moving.GetSpacing())[None, :]
    p2i = np.linalg.inv(i2p)
    c = (y - np.asarray(moving.GetOrigin())) @ p2i.T
    mz, my, mx = M.shape
    size = np.array([mx, my, mz], np.float64)
    keep &= np.all((c >= -0.5) & (c < size - 0.5), axis=1)
    if moving_mask is not None:
        r = np.floor(np.where(keep[:, None], c, 0.0) + 0.5).astype(np.int64)
        keep &= moving_mask.array[r[:, 2], r[:, 1], r[:, 0]] != 0
    c, x, fv = c[keep], x[keep], F.reshape(-1)[q][keep]
    cc = np.maximum(c, 0.0)                      # LinearInterpolateImageFunction: base clamped up to 0, distance 0 there
    b = np.floor(cc).astype(np.int64)
    d = cc - b
    u = np.minimum(b + 1, (size - 1).astype(np.int64))
    b = np.minimum(b, (size - 1).astype(np.int64))
    g = lambda zz, yy, xx: M[zz, yy, xx]
    v000, v100 = g(b[:, 2], b[:, 1], b[:, 0]), g(b[:, 2], b[:, 1], u[:, 0])
    v010, v110 = g(b[:, 2], u[:, 1], b[:, 0]), g(b[:, 2], u[:, 1], u[:, 0])
    v001, v101 = g(u[:, 2], b[:, 1], b[:, 0]), g(u[:, 2], b[:, 1], u[:, 0])
    v011, v111 = g(u[:, 2], u[:, 1], b[:, 0]), g(u[:, 2], u[:, 1], u[:, 0])
    d0, d1, d2 = d[:, 0], d[:, 1], d[:, 2]
    a00, a10, a01, a11 = v100 - v000, v110 - v010, v101 - v001, v111 - v011
    vx00, vx10, vx01, vx11 = v000 + a00 * d0, v010 + a10 * d0, v001 + a01 * d0, v011 + a11 * d0
    vxx0, vxx1 = vx00 + (vx10 - vx00) * d1, vx01 + (vx11 - vx01) * d1
    mval = vxx0 + (vxx1 - vxx0) * d2
    gx0, gx1 = a00 + (a10 - a00) * d1, a01 + (a11 - a01) * d1
    gi = np.stack([gx0 + (gx1 - gx0) * d2, (vx10 - vx00) + ((vx11 - vx01) - (vx10 - vx00)) * d2, vxx1 - vxx0], axis=1)
    gy = gi @ p2i                                 # d/dy_j = sum_i g_i P2I[i][j]
    h = gy @ np.asarray(initial_matrix, np.float64).reshape(3, 3)   # A_i^T gy, row-vector form
    if kind == "raw":
        return fv, mval, h, x - np.asarray(center, np.float64)
    if kind == "correlation":
        xc = x - np.asarray(center, np.float64)
        out = np.zeros(42)
        out[0:6] = [float(fv.size), fv.sum(), mval.sum(), (fv * fv).sum(), (mval * mval).sum(), (fv * mval).sum()]
        for k, wt in enumerate((np.ones_like(fv), fv, mval)):
            wh = wt[:, None] * h
            out[6 + 12 * k: 9 + 12 * k] = wh.sum(axis=0)
            out[9 + 12 * k: 18 + 12 * k] = (wh[:, :, None] * xc[:, None, :]).sum(axis=0).reshape(9)
        return out
    dd = mval - fv
    w = 2.0 * dd[:, None] * h
    out = np.zeros(14)
    out[0], out[1] = float((dd * dd).sum()), float(dd.size)
    out[2:5] = w.sum(axis=0)
    out[5:14] = (w[:, :, None] * (x - np.asarray(center, np.float64))[:, None, :]).sum(axis=0).reshape(9)
    return out


# ---- label utilities (utils/crop.py:24-99, label/utils.py:23-58, multiatlas/run.py:387-437) -----------------------------
def label_to_roi(label, expansion_mm=[0, 0, 0], return_as_list=False):
    """crop.py:24-71 with LabelStatisticsImageFilter.GetBoundingBox restated in numpy."""
    if isinstance(label, (list, tuple)):
        ref_lab = sum(l.array.astype(np.float64) for l in label) > 0
        spacing, full = np.array(label[0].GetSpacing()), np.array(label[0].GetSize())
    else:
        ref_lab = label.array > 0
        spacing, full = np.array(label.GetSpacing()), np.array(label.GetSize())
    zz, yy, xx = np.nonzero(ref_lab)
    bounding_box = [xx.min(), xx.max(), yy.min(), yy.max(), zz.min(), zz.max()]
    index = [bounding_box[x * 2] for x in range(3)]
    size = [bounding_box[(x * 2) + 1] - bounding_box[x * 2] + 1 for x in range(3)]
    expansion = (np.array(expansion_mm) / spacing).astype(int)
    crop_box_index = np.max([index - expansion, np.array([0, 0, 0])], axis=0)
    crop_box_size = np.min([full - crop_box_index, np.array(size) + 2 * expansion], axis=0)
    crop_box_size = [int(i) for i in crop_box_size]
    crop_box_index = [int(i) for i in crop_box_index]
    if return_as_list:
        return crop_box_index + crop_box_size
    return crop_box_size, crop_box_index


def crop_to_roi(image, size, index):
    """sitk.RegionOfInterest: array slice, origin moved to the first voxel kept."""
    x, y, z = index
    sx, sy, sz = size
    arr = image.array[z:z + sz, y:y + sy, x:x + sx].copy()
    d = np.asarray(image.GetDirection(), np.float64).reshape(3, 3)
    origin = np.asarray(image.GetOrigin()) + d @ (np.asarray(image.GetSpacing()) * np.asarray(index, np.float64))
    return Image(arr, image.GetSpacing(), tuple(origin), image.GetDirection())


def paste(destination_image, source_image, source_size, source_index, destination_index):
    out = destination_image.array.copy()
    sx, sy, sz = source_size
    x, y, z = source_index
    dx, dy, dz = destination_index
    out[dz:dz + sz, dy:dy + sy, dx:dx + sx] = source_image.array[z:z + sz, y:y + sy, x:x + sx]
    return _like(out, destination_image)


def correct_volume_overlap(binary_label_dict, assign_overlap_to_largest=True):
    """label/utils.py:23-58 (prime encoding restated as set logic: a voxel stays with the first structure, in volume rank
    order, that contains it)."""
    keys = list(binary_label_dict.keys())
    vals = [int(binary_label_dict[k].array.sum()) for k in keys]
    rank = np.argsort(vals)[::-1] if assign_overlap_to_largest else np.argsort(vals)
    ranked = [keys[i] for i in rank]
    combined = sum((binary_label_dict[k].array > 0).astype(np.int64) for k in keys) > 0
    out = {}
    for k in ranked:
        o = combined & (binary_label_dict[k].array > 0)
        out[k] = _like(o.astype(np.uint8), binary_label_dict[k])
        combined = combined & ~o
    return out


def binary_morphological_closing(image, kernel_radius, structure):
    """sitk.BinaryMorphologicalClosing (SafeBorder on): scipy dilation then erosion on a grid padded with background;
    ``structure`` is the [z, y, x] boolean structuring element."""
    import scipy.ndimage as ndi

    r = [int(v) for v in kernel_radius]
    a = np.pad(image.array > 0, ((r[2], r[2]), (r[1], r[1]), (r[0], r[0])))
    d = ndi.binary_dilation(a, structure=structure)
    e = ndi.binary_erosion(d, structure=structure, border_value=0)
    e = e[r[2]:e.shape[0] - r[2], r[1]:e.shape[1] - r[1], r[0]:e.shape[2] - r[0]]
    return _like(e.astype(np.uint8), image)


def exponentiate_field(velocity_field, number_of_iterations=None, maximum_number_of_iterations=20):
    """Scaling and squaring as ExponentialDisplacementFieldImageFilter does it: u_0 = v / 2^N, u <- u + Resample(u, DisplacementFieldTransform(u))
    N times; N automatic = smallest N with max|v| / 2^N <= half the smallest spacing (restates platipy_b200.registration.exponentiate_field)."""
    v = velocity_field
    if number_of_iterations is None:
        max_norm = float(np.sqrt((v.array ** 2).sum(axis=-1)).max())
        half_voxel = 0.5 * min(v.GetSpacing())
        n = 0
        while max_norm / (2.0 ** n) > half_voxel and n < int(maximum_number_of_iterations):
            n += 1
    else:
        n = int(number_of_iterations)
    u = _like(v.array / (2.0 ** n) if n > 0 else v.array.copy(), v, True)
    for _ in range(n):
        warped = resample(u, u, sk.DisplacementFieldTransform(u))
        u = _like(u.array + warped.array, u, True)
    return u
