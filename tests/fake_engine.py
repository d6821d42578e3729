"""TEST INFRASTRUCTURE: a stand-in for ``platipy_b200.engine.Engine`` whose methods are answered by the oracle (and, for the
windowed correlation, by the kernel source under the host emulation).  It exists for one purpose: the build container has no GPU,
so host-side glue written there (argument order, composition order, dtypes, representation handling in comparison.py, the
patch-correlation vote, the metric / optimiser plumbing of linear_registration ...) would otherwise not execute even once before
the GPU tests run it on a B200.  ``tests/test_host_glue_on_fake_engine.py`` runs the bodies of the newest GPU tests against this
stand-in.  It proves nothing about the kernels (those are checked under tests/emu and on the GPU) and nothing in the package can
reach it: it is installed by monkeypatching ``Engine.get`` inside a test.

"Device" images are ``DeviceImage`` handles around CPU torch tensors; ``stream`` is None, which makes ``torch.cuda.stream(...)`` a
no-op context."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from oracle import itk_oracle as orc
from oracle import platipy_ref as ref
from platipy_b200 import sitk_compat as sk
from platipy_b200.engine import _DT, DeviceImage
from platipy_b200.sitk_compat import Image


def _arr(d):
    """numpy view of a scalar device image, or the AoS [z, y, x, 3] copy of a vector one."""
    a = d.tensor.numpy()
    if a.dtype != d.np_dtype:
        a = a.view(d.np_dtype)
    return np.moveaxis(a, 0, -1) if d.is_vector else a


def _img(d):
    return Image(np.ascontiguousarray(_arr(d)), d.GetSpacing(), d.GetOrigin(), d.GetDirection(), d.is_vector)


class _Grid:
    def __init__(self, src):
        self.size, self.spacing, self.origin, self.direction = src.GetSize(), src.GetSpacing(), src.GetOrigin(), src.GetDirection()


class FakeEngine:
    stream = None
    device = torch.device("cpu")

    def __init__(self, emu=None):
        self.emu = emu
        self.calls = {}

    def _note(self, name):
        self.calls[name] = self.calls.get(name, 0) + 1

    # -- plumbing -------------------------------------------------------------------------------------------------------
    def _wrap(self, arr, like, is_vector=False):
        arr = np.ascontiguousarray(arr)
        if is_vector:
            arr = np.ascontiguousarray(np.moveaxis(arr, -1, 0))
        return DeviceImage(torch.from_numpy(arr.copy()), arr.dtype, like.GetSpacing(), like.GetOrigin(), like.GetDirection(), is_vector)

    def empty(self, shape, np_dtype):
        return torch.zeros(tuple(int(s) for s in shape), dtype=_DT[np.dtype(np_dtype)][1])

    zeros = empty

    def synchronize(self):
        pass

    wait_caller = release_to_caller = synchronize

    def launch_count(self):
        return 0

    def to_device(self, image):
        if isinstance(image, DeviceImage):
            return image
        image = sk.to_native(image)
        return self._wrap(image.array, image, image.is_vector)

    def to_host(self, dimg, pinned=True):
        return _img(dimg)

    class _Done:
        def synchronize(self):
            pass

        def query(self):
            return True

    class _Stream:
        def wait_event(self, ev):
            pass

    def to_device_async(self, image):
        return self.to_device(image), None

    def to_host_async(self, dimg):
        return _img(dimg), FakeEngine._Done()

    # -- elementwise ----------------------------------------------------------------------------------------------------
    def cast(self, d, np_dtype):
        self._note("cast")
        np_dtype = np.dtype(np_dtype)
        if np_dtype == d.np_dtype:
            return d
        return self._wrap(sk.Cast(_img(d), sk.dtype_to_pixel_id(np_dtype, d.is_vector)).array, d, d.is_vector)

    def minmax(self, d):
        self._note("minmax")
        a = _arr(d)
        return float(a.min()), float(a.max())

    def binary_threshold(self, d, lower, upper=255.0):
        self._note("binary_threshold")
        a = _arr(d).astype(np.float64)
        return self._wrap(((a >= lower) & (a <= upper)).astype(np.uint8), d)

    def u8_binary_op(self, a, b, op):
        self._note("u8_binary_op")
        fn = (np.bitwise_or, np.bitwise_and, np.add, np.bitwise_xor)[int(op)]
        return self._wrap(fn(_arr(a), _arr(b)), a)

    def mask_image(self, d, mask, outside_value=0.0):
        self._note("mask_image")
        a, m = _arr(d), _arr(mask) != 0
        if d.is_vector:
            m = m[..., None]
        return self._wrap(np.where(m, a, a.dtype.type(outside_value)), d, d.is_vector)

    def divide_scalar(self, d, divisor):
        self._note("divide_scalar")
        a = _arr(d)
        return self._wrap(a / a.dtype.type(divisor), d)

    def scale_shift(self, d, mul=1.0, add=0.0, take_abs=False):
        self._note("scale_shift")
        if d.np_dtype not in (np.dtype(np.float32), np.dtype(np.float64)) or d.is_vector:
            raise TypeError("image arithmetic on the device is implemented for scalar Float32 / Float64 images")
        a = _arr(d)
        v = np.abs(a) if take_abs else a
        return self._wrap(v * a.dtype.type(mul) + a.dtype.type(add), d)

    # -- distance maps, contours, morphology ------------------------------------------------------------------------------
    def signed_maurer_distance_map(self, mask, inside_is_positive=False, squared_distance=False, use_image_spacing=True):
        self._note("signed_maurer_distance_map")
        return self._wrap(orc.signed_maurer_distance_map(_arr(mask), mask.GetSpacing(), inside_is_positive, squared_distance, use_image_spacing), mask)

    def label_contour(self, mask, fully_connected=False):
        self._note("label_contour")
        return self._wrap(orc.label_contour(_arr(mask), fully_connected), mask)

    def label_contour_slicewise(self, mask):
        self._note("label_contour_slicewise")
        a = _arr(mask)
        return self._wrap(np.stack([orc.label_contour(a[i:i + 1], False)[0] for i in range(a.shape[0])]), mask)

    def binary_dilate(self, mask, offsets, boundary_to_foreground=False):
        self._note("binary_dilate")
        return self._wrap(orc.binary_morph(_arr(mask), offsets, True, boundary_to_foreground), mask)

    def binary_erode(self, mask, offsets, boundary_to_foreground=True):
        self._note("binary_erode")
        return self._wrap(orc.binary_morph(_arr(mask), offsets, False, boundary_to_foreground), mask)

    def binary_closing(self, mask, radius, offsets):
        self._note("binary_closing")
        r = [int(v) for v in radius]
        offs = np.asarray(offsets).reshape(-1, 3)
        st = np.zeros((2 * r[2] + 1, 2 * r[1] + 1, 2 * r[0] + 1), bool)
        st[offs[:, 2] + r[2], offs[:, 1] + r[1], offs[:, 0] + r[0]] = True
        return self._wrap(ref.binary_morphological_closing(_img(mask), r, st).array, mask)

    # -- label utilities ------------------------------------------------------------------------------------------------
    def bounding_box(self, mask):
        self._note("bounding_box")
        zz, yy, xx = np.nonzero(_arr(mask))
        if zz.size == 0:
            return [2 ** 31 - 1] * 3 + [-1] * 3
        return [int(xx.min()), int(yy.min()), int(zz.min()), int(xx.max()), int(yy.max()), int(zz.max())]

    def region_copy(self, src, src_index, dst, dst_index, region_size):
        self._note("region_copy")
        (sx, sy, sz), (dx, dy, dz), (nx, ny, nz) = src_index, dst_index, region_size
        for lo, n, size in zip(tuple(src_index) + tuple(dst_index), tuple(region_size) * 2, src.GetSize() + dst.GetSize()):
            if lo < 0 or n <= 0 or lo + n > size:
                raise RuntimeError("requested region is (at least partially) outside the largest possible region")
        dst.tensor[dz:dz + nz, dy:dy + ny, dx:dx + nx] = src.tensor[sz:sz + nz, sy:sy + ny, sx:sx + nx]
        return dst

    # -- resampling / smoothing -------------------------------------------------------------------------------------------
    def resample(self, d, out_geom_src=None, transform=None, interpolator=sk.sitkLinear, default_value=0.0):
        self._note("resample")
        g = _Grid(out_geom_src if out_geom_src is not None else d)
        if d.is_vector:
            return self.resample_vec3(d, out_geom_src if out_geom_src is not None else d, transform, default_value)
        transform = self._host_transform(transform)  # displacement-field transforms made on the "device" carry a DeviceImage
        out = ref.resample(_img(d), None, transform, interpolator, default_value, g.size, g.spacing, g.origin, g.direction)
        return DeviceImage(torch.from_numpy(np.ascontiguousarray(out.array)), out.array.dtype, g.spacing, g.origin, g.direction, False)

    def discrete_gaussian(self, d, variance, maximum_kernel_width=32, maximum_error=0.01, use_image_spacing=True):
        self._note("discrete_gaussian")
        var = [float(variance)] * 3 if np.isscalar(variance) else [float(v) for v in variance]
        return self._wrap(orc.discrete_gaussian_f32(_arr(d), orc.geom_of(_img(d)), var, maximum_kernel_width, maximum_error, use_image_spacing), d)

    def smooth_and_resample(self, d, variance, maximum_kernel_width, out_geom_src, interpolator=sk.sitkLinear, allow_restricted=True):
        self._note("smooth_and_resample")
        return self.resample(self.discrete_gaussian(d, variance, maximum_kernel_width), out_geom_src, None, interpolator, 0.0)

    # -- reductions behind linear_registration ----------------------------------------------------------------------------
    def image_moments(self, d):
        self._note("image_moments")
        return ref.image_moments(_img(d))

    @staticmethod
    def _opt(d):
        return _img(d) if d is not None else None

    def linreg_meansq(self, fixed, moving, total_matrix, total_offset, initial_matrix, center, fixed_mask=None, moving_mask=None, stride=1):
        self._note("linreg_meansq")
        return ref.linreg_meansq(_img(fixed), _img(moving), total_matrix, total_offset, initial_matrix, center, self._opt(fixed_mask), self._opt(moving_mask), stride)

    def linreg_correlation(self, fixed, moving, total_matrix, total_offset, initial_matrix, center, fixed_mask=None, moving_mask=None, stride=1):
        self._note("linreg_correlation")
        return ref.linreg_correlation(_img(fixed), _img(moving), total_matrix, total_offset, initial_matrix, center, self._opt(fixed_mask),
                                      self._opt(moving_mask), stride)

    def linreg_mattes_histogram(self, fixed, moving, total_matrix, total_offset, fixed_bins, moving_bins, n_bins=50, fixed_mask=None, moving_mask=None,
                                stride=1):
        self._note("linreg_mattes_histogram")
        return ref.linreg_mattes(_img(fixed), _img(moving), total_matrix, total_offset, np.eye(3), np.zeros(3), fixed_bins, moving_bins, n_bins, None,
                                 self._opt(fixed_mask), self._opt(moving_mask), stride)

    def linreg_mattes_derivative(self, fixed, moving, total_matrix, total_offset, initial_matrix, center, fixed_bins, moving_bins, table, fixed_mask=None,
                                 moving_mask=None, stride=1):
        self._note("linreg_mattes_derivative")
        return ref.linreg_mattes(_img(fixed), _img(moving), total_matrix, total_offset, initial_matrix, center, fixed_bins, moving_bins, np.asarray(table).shape[0],
                                 table, self._opt(fixed_mask), self._opt(moving_mask), stride)[2]

    # -- field templates, recursive Gaussian, Demons (generation.py) -------------------------------------------------------------
    def constant_field(self, grid, vector, mask=None):
        self._note("constant_field")
        x, y, z = grid.GetSize()
        arr = np.zeros((z, y, x, 3)) + np.asarray(vector, dtype=np.float64)
        if mask is not None:
            arr = np.where((_arr(mask) != 0)[..., None], arr, 0.0)
        return self._wrap(arr, grid, True)

    def radial_bend_field(self, mask, reference_index, axis, scale, clip_axis=-1, clip_keep_upper=True):
        self._note("radial_bend_field")
        body = _arr(mask) != 0
        zz, yy, xx = np.nonzero(np.ones_like(body))
        idx = np.stack([xx, yy, zz], axis=1).reshape(body.shape + (3,))
        if clip_axis >= 0:
            c = idx[..., clip_axis]
            body = body & ((c >= reference_index[clip_axis]) if clip_keep_upper else (c < reference_index[clip_axis]))
        rel = (idx - np.asarray(reference_index)).astype(np.int64)
        arr = np.where(body[..., None], np.cross(rel, np.asarray(axis, dtype=np.float64)) * float(scale), 0.0)
        return self._wrap(arr, mask, True)

    def recursive_gaussian(self, dfield, sigma):
        self._note("recursive_gaussian")
        out = orc.recursive_gaussian_vec3(_arr(dfield), orc.geom_of(_img(dfield)), [float(v) for v in sigma])
        dfield.tensor.copy_(torch.from_numpy(np.ascontiguousarray(np.moveaxis(out, -1, 0))))  # the entry point works in place
        return dfield

    def multiscale_demons(self, fixed, moving, cfg, initial_field=None, initial_on_fixed_grid=False):
        self._note("multiscale_demons")
        n = cfg.n_levels
        flt = ref.DemonsFilter()
        flt.SetStandardDeviations(list(cfg.demons.std_dev))
        flt.SetSmoothDisplacementField(bool(cfg.demons.smooth_displacement_field))
        flt.SetSmoothUpdateField(bool(cfg.demons.smooth_update_field))
        stats = []
        dvf = ref.multiscale_demons(flt, _img(fixed), _img(moving), initial_displacement_field=_img(initial_field) if initial_field is not None else None,
                                    isotropic_resample=bool(cfg.isotropic_resample), resolution_staging=list(cfg.resolution_staging[:n]),
                                    smoothing_sigmas=list(cfg.smoothing_sigmas[:n]), iteration_staging=list(cfg.iteration_staging[:n]),
                                    interp_order=int(cfg.interp_order), level_stats=stats)
        for st in stats:
            st["gpu_ms"] = 0.0
        return self._wrap(dvf.array, fixed, True), stats

    # -- fusion pieces used by iterative atlas removal -----------------------------------------------------------------------------
    def process_probability(self, d, threshold=0.5):
        self._note("process_probability")
        return self._wrap(ref.process_probability_image(_img(d), threshold).array, d)

    def vote_accumulate(self, label, weight, num, den, first):
        # real float32 arithmetic on the accumulators (so that sharded runs can exchange them), as vote_accumulate_kernel does it
        self._note("vote_accumulate")
        t = _arr(weight).astype(np.float32) * _arr(label).astype(np.float32)
        w = _arr(weight).astype(np.float32)
        if first:
            num.copy_(torch.from_numpy(t))
            if den is not None:
                den.copy_(torch.from_numpy(w))
        else:
            num.copy_(torch.from_numpy(num.numpy() + t))
            if den is not None:
                den.copy_(torch.from_numpy(den.numpy() + w))

    def vote_finalize(self, num, den, geom_src, smooth_variance, threshold):
        # fusion.py:264-288 from the accumulators, restating orc_combine_labels_f32's tail step for step
        self._note("vote_finalize")
        comb = num.numpy().astype(np.float32)
        if den is not None:
            d = den.numpy().astype(np.float32).copy()
            d[d == 0] = 1.0
            comb = comb / d
        sm = orc.discrete_gaussian_f32(np.ascontiguousarray(comb), orc.geom_of(geom_src), [smooth_variance] * 3, 32, 0.01, True)
        mn, mx = np.float32(sm.min()), np.float32(sm.max())
        if abs(np.float32(mx - mn)) > np.finfo(np.float32).eps:
            scale = 1.0 / (float(mx) - float(mn))
        elif float(mx) != 0.0:
            scale = 1.0 / float(mx)
        else:
            scale = 0.0
        shift = 0.0 - float(mn) * scale
        r = (sm.astype(np.float64) * scale + shift).astype(np.float32)
        r = np.where(r > 1, np.float32(1), r)
        r = np.where(r < 0, np.float32(0), r)
        if threshold:
            r = np.where((r >= np.float32(threshold)) & (r <= 1), r, np.float32(0))
        out = np.ascontiguousarray(r, dtype=np.float32)
        return DeviceImage(torch.from_numpy(out), np.float32, geom_src.GetSpacing(), geom_src.GetOrigin(), geom_src.GetDirection(), False)

    # -- the remaining entry points of the atlas pipeline (multiatlas.py, fusion.py, registration.py) -----------------------------------
    def _host_transform(self, transform):
        """sk transform whose displacement fields are host images (device-made transforms carry a DeviceImage)."""
        if transform is None:
            return None
        parts = []
        for t in reversed(transform.flatten()):
            if isinstance(t, sk.DisplacementFieldTransform):
                f = t._device_cache[1] if t._device_cache is not None else t.GetDisplacementField()
                parts.append(sk.DisplacementFieldTransform(_img(f) if isinstance(f, DeviceImage) else f))
            else:
                parts.append(t)
        return sk.CompositeTransform(parts)

    def resample_batch(self, images, out_geom_src, transform, interpolators, default_values):
        self._note("resample_batch")
        return [self.resample(im, out_geom_src, transform, ip, dv) for im, ip, dv in zip(images, interpolators, default_values)]

    def resample_vec3(self, dfield, out_geom_src, transform=None, default_value=0.0):
        self._note("resample_vec3")
        g = _Grid(out_geom_src)
        out = ref.resample(_img(dfield), None, self._host_transform(transform), sk.sitkLinear, default_value, g.size, g.spacing, g.origin, g.direction)
        return DeviceImage(torch.from_numpy(np.ascontiguousarray(np.moveaxis(out.array, -1, 0))), np.float64, g.spacing, g.origin, g.direction, True)

    def transform_to_dvf(self, transform, grid):
        self._note("transform_to_dvf")
        arr = orc.transform_to_dvf(orc.geom_of(grid), ref._chain_of(self._host_transform(transform)))
        return self._wrap(arr, grid, True)

    def demons_execute(self, fixed, moving, params):
        self._note("demons_execute")
        p = orc.demons_params(list(params.std_dev), params.number_of_iterations, list(params.update_std_dev), bool(params.smooth_displacement_field),
                              bool(params.smooth_update_field), params.max_error, params.max_kernel_width, params.max_rms_error,
                              params.max_update_step_length, params.intensity_difference_threshold, params.denominator_threshold)
        f, m = _img(fixed), _img(moving)
        D, st = orc.demons_execute(f.array, orc.geom_of(f), m.array, orc.geom_of(m), p)
        st.update(gpu_ms=0.0, voxels=fixed.GetNumberOfPixels())
        return self._wrap(D, fixed, True), st

    def weight_map(self, target, moving, vote_type, factor=1e12, sigma=2.0, epsilon=1e-5):
        self._note("weight_map")
        name = ("unweighted", "global", "local")[int(vote_type)]
        params = {"factor": factor, "sigma": sigma, "epsilon": epsilon, "normalise": False}
        return self._wrap(ref.compute_weight_map(_img(target), _img(moving), name, params).array, target)

    def weight_map_block(self, target, moving, radius, factor, gain):
        self._note("weight_map_block")
        params = {"factor": factor, "gain": gain, "blockSize": tuple(int(v) for v in radius), "normalise": False}
        return self._wrap(ref.compute_weight_map(_img(target), _img(moving), "block", params).array, target)

    def normalise_by_max(self, weight, mask=None):
        self._note("normalise_by_max")
        out = ref._normalise(_arr(weight).copy(), True if mask is None else _img(mask))
        weight.tensor.copy_(torch.from_numpy(np.ascontiguousarray(out)))
        return weight

    def pack_decision(self, label, bit, packed, first):
        self._note("pack_decision")
        if first:
            packed.zero_()
        packed += (label.tensor != 0).to(packed.dtype) << int(bit)

    def unpack_decision(self, packed, bit, like):
        self._note("unpack_decision")
        return like.like(((packed >> int(bit)) & 1).to(torch.uint8), np.uint8, False)

    # -- compact exchange formats of the sharded fusion ---------------------------------------------------------------------------------
    def pack_label(self, label, bit, packed, first):
        self._note("pack_label")
        if first:
            packed.zero_()
        packed |= (label.tensor != 0).to(packed.dtype) << int(bit)

    def staple_packed(self, packed, holder_mask, like, confidence_weight=1.0, max_iterations=0xFFFFFFFF, threshold=1e-4, rescale=True, want_info=False):
        self._note("staple_packed")
        bits = [b for b in range(32) if (int(holder_mask) >> b) & 1]
        raw = packed.numpy().astype(np.int64) & ((1 << 32) - 1 if packed.element_size() == 4 else (1 << (8 * packed.element_size())) - 1)
        W, p, q, it = orc.staple([((raw >> b) & 1).astype(np.uint8) for b in bits], confidence_weight, max_iterations)
        if rescale or threshold:
            W = orc.rescale_threshold_f64(W, threshold)
        res = self._wrap(W, like)
        return (res, {"p": list(p), "q": list(q), "elapsed_iterations": int(it)}) if want_info else res

    def count_accumulate(self, label, counts, first, flag):
        self._note("count_accumulate")
        if first:
            counts.zero_()
        counts += label.tensor
        if bool((label.tensor > 1).any()):
            flag |= 1

    def vote_finalize_counts(self, counts, n_holders, geom_src, smooth_variance, threshold):
        self._note("vote_finalize_counts")
        # n_holders binary labels whose sum is the count volume, each with weight 1: the same float32 num / den as the real votes
        c = counts.numpy()
        labels = [(c > k).astype(np.uint8) for k in range(max(int(n_holders), 1))]
        ones = np.ones(c.shape, np.float32)
        out = orc.combine_labels_f32(labels, [ones] * len(labels), orc.geom_of(geom_src), smooth_variance, threshold)
        return DeviceImage(torch.from_numpy(out), np.float32, geom_src.GetSpacing(), geom_src.GetOrigin(), geom_src.GetDirection(), False)

    def staple(self, decisions, confidence_weight=1.0, max_iterations=0xFFFFFFFF, threshold=1e-4, rescale=True):
        self._note("staple")
        W, p, q, it = orc.staple([_arr(d) for d in decisions], confidence_weight, max_iterations)
        if rescale or threshold:
            W = orc.rescale_threshold_f64(W, threshold)
        return self._wrap(W, decisions[0]), {"p": list(p), "q": list(q), "elapsed_iterations": int(it)}

    def binary_fillhole(self, d, fully_connected=False):
        self._note("binary_fillhole")
        return self._wrap(orc.binary_fillhole(_arr(d)), d)

    def largest_component(self, d, fully_connected=False, want_info=False):
        self._note("largest_component")
        out, ncomp, nvox = orc.largest_component(_arr(d))
        res = self._wrap(out, d)
        return (res, {"n_components": ncomp, "voxels": nvox}) if want_info else res

    def resolve_overlap(self, labels_ranked):
        self._note("resolve_overlap")
        taken = np.zeros(_arr(labels_ranked[0]).shape, bool)
        outs = []
        for l in labels_ranked:
            has = _arr(l) != 0
            outs.append(self._wrap((has & ~taken).astype(np.uint8), l))
            taken |= has
        return outs

    # -- patch correlation: the kernel source itself, under the host emulation ------------------------------------------------
    def patch_correlation(self, target, moving, window):
        self._note("patch_correlation")
        t, m = np.ascontiguousarray(_arr(target), np.float32), np.ascontiguousarray(_arr(moving), np.float32)
        nz, ny, nx = t.shape
        out = np.empty(t.shape, np.float64)
        P = lambda a: a.ctypes.data_as(C.c_void_p)
        self.emu.emu_patch_correlation(P(t), P(m), nx, ny, nz, int(window[0]), int(window[1]), int(window[2]), P(out), C.c_uint(4), C.c_uint(64))
        return self._wrap(out, target)
