"""
Device plumbing: one ``Engine`` per GPU owns a CUDA stream (a ``torch.cuda.Stream``), the ``b200reg`` context
bound to it, and wraps every C-ABI entry point on ``DeviceImage`` handles.

PyTorch is used only for device/pinned memory, the stream and (in ``platipy_b200.multiatlas``) the NCCL
process group; all arithmetic is done by the CUDA kernels of ``libb200reg.so``.
"""
from __future__ import annotations

import ctypes as C
import threading
import weakref

import numpy as np
import torch

from . import _abi
from . import sitk_compat as sk
from .sitk_compat import Image

# numpy dtype -> (b200reg dtype id, torch container dtype of the same width)
_DT = {
    np.dtype(np.int8): (0, torch.int8), np.dtype(np.uint8): (1, torch.uint8),
    np.dtype(np.int16): (2, torch.int16), np.dtype(np.uint16): (3, torch.int16),
    np.dtype(np.int32): (4, torch.int32), np.dtype(np.uint32): (5, torch.int32),
    np.dtype(np.int64): (6, torch.int64), np.dtype(np.uint64): (7, torch.int64),
    np.dtype(np.float32): (8, torch.float32), np.dtype(np.float64): (9, torch.float64),
}
_MASK_DT = {1: 1, 2: 3, 4: 5}  # element size of a decision mask -> b200reg dtype id (UInt8, UInt16, UInt32)
_SIGNED_VIEW = {np.dtype(np.uint16): np.int16, np.dtype(np.uint32): np.int32, np.dtype(np.uint64): np.int64}


def _as_torch_host(arr):
    """numpy array -> torch CPU tensor sharing memory (unsigned types travel in a signed container)."""
    arr = np.ascontiguousarray(arr)
    if arr.dtype in _SIGNED_VIEW:
        arr = arr.view(_SIGNED_VIEW[arr.dtype])
    return torch.from_numpy(arr)


class DeviceImage:
    """A volume resident in HBM.  Scalar: tensor ``[z, y, x]``.  Vector (f64 x 3): SoA tensor ``[3, z, y, x]``."""

    __slots__ = ("tensor", "np_dtype", "spacing", "origin", "direction", "is_vector")

    def __init__(self, tensor, np_dtype, spacing, origin, direction, is_vector=False):
        self.tensor = tensor
        self.np_dtype = np.dtype(np_dtype)
        self.spacing = tuple(float(s) for s in spacing)
        self.origin = tuple(float(s) for s in origin)
        self.direction = tuple(float(s) for s in direction)
        self.is_vector = bool(is_vector)

    def GetSize(self):
        z, y, x = self.tensor.shape[-3:]
        return (int(x), int(y), int(z))

    def GetSpacing(self):
        return self.spacing

    def GetOrigin(self):
        return self.origin

    def GetDirection(self):
        return self.direction

    def GetPixelID(self):
        return sk.dtype_to_pixel_id(self.np_dtype, self.is_vector)

    def GetNumberOfPixels(self):
        x, y, z = self.GetSize()
        return x * y * z

    @property
    def ptr(self):
        return C.c_void_p(self.tensor.data_ptr())

    @property
    def geom(self):
        return _abi.make_geom(self.GetSize(), self.spacing, self.origin, self.direction)

    @property
    def dtype_id(self):
        return _DT[self.np_dtype][0]

    def like(self, tensor, np_dtype=None, is_vector=None):
        return DeviceImage(tensor, self.np_dtype if np_dtype is None else np_dtype, self.spacing, self.origin, self.direction,
                           self.is_vector if is_vector is None else is_vector)

    # image-with-constant arithmetic (SimpleITK's Image operators), as far as a ``correlation_function`` needs it
    # (fusion.py:138-146: ``lambda x: x + 1``, ``abs``); each is one kernel on the engine's stream
    def _affine(self, mul, add, take_abs=False):
        if not np.isscalar(mul) or not np.isscalar(add):
            raise TypeError("DeviceImage arithmetic supports scalar constants only")
        return Engine.get(self.tensor.device.index).scale_shift(self, mul, add, take_abs)

    def __add__(self, c):
        return self._affine(1.0, c)

    __radd__ = __add__

    def __sub__(self, c):
        return self._affine(1.0, -c) if np.isscalar(c) else NotImplemented

    def __rsub__(self, c):
        return self._affine(-1.0, c)

    def __mul__(self, c):
        return self._affine(c, 0.0)

    __rmul__ = __mul__

    def __truediv__(self, c):
        if not np.isscalar(c):
            raise TypeError("DeviceImage arithmetic supports scalar constants only")
        return Engine.get(self.tensor.device.index).divide_scalar(self, c)

    def __neg__(self):
        return self._affine(-1.0, 0.0)

    def __abs__(self):
        return self._affine(1.0, 0.0, True)


def pinned_empty(shape, dtype):
    """numpy array backed by pinned host memory (keeps the owning tensor alive through ``.base``)."""
    dtype = np.dtype(dtype)
    container = _DT[dtype][1]
    t = torch.empty(tuple(int(s) for s in shape), dtype=container, pin_memory=True)
    a = t.numpy()
    if dtype in _SIGNED_VIEW:
        a = a.view(dtype)
    return a


class _PinnedPool:
    """Pinned host buffers of the asynchronous downloads, recycled.  ``cudaHostAlloc`` of a 1.6 GB field buffer takes about a second
    (longer when eight ranks pin memory at once), so a buffer goes back to this pool -- not to the allocator -- when the numpy array
    that was handed to the caller, and every view of it, has been garbage collected (a ``weakref.finalize`` on the base array)."""

    MAX_PER_KEY = 4

    def __init__(self):
        self.free = {}
        self.lock = threading.Lock()

    def take(self, shape, torch_dtype):
        key = (tuple(int(v) for v in shape), torch_dtype)
        with self.lock:
            lst = self.free.get(key)
            if lst:
                return lst.pop()
        return torch.empty(key[0], dtype=torch_dtype, pin_memory=True)

    def give(self, tensor):
        key = (tuple(tensor.shape), tensor.dtype)
        with self.lock:
            lst = self.free.setdefault(key, [])
            if len(lst) < self.MAX_PER_KEY:
                lst.append(tensor)


def pinned_image(image):
    """Copy of ``image`` whose pixel buffer lives in pinned host memory (fast H2D)."""
    image = sk.to_native(image)
    buf = pinned_empty(image.array.shape, image.array.dtype)
    np.copyto(buf, image.array)
    out = Image.__new__(Image)
    out._arr = buf
    out._spacing, out._origin, out._direction, out._is_vector = image._spacing, image._origin, image._direction, image._is_vector
    return out


class Engine:
    """Per-GPU engine; ``Engine.get()`` returns the singleton of the current (or given) CUDA device."""

    _instances = {}
    _lock = threading.Lock()

    @classmethod
    def get(cls, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("platipy_b200 needs a CUDA device (sm_100a); no CPU fallback exists")
        if device is None:
            device = torch.cuda.current_device()
        device = int(torch.device("cuda", device).index if not isinstance(device, int) else device)
        with cls._lock:
            if device not in cls._instances:
                cls._instances[device] = cls(device)
            return cls._instances[device]

    def __init__(self, device):
        self.lib = _abi.load()
        self.device = torch.device("cuda", device)
        with torch.cuda.device(self.device):
            self.stream = torch.cuda.Stream(device=self.device)
            # copy streams of the pipelined host API (to_device_async / to_host_async): PCIe transfers in both directions ride
            # beside the compute stream instead of in front of / behind it
            self.copy_in = torch.cuda.Stream(device=self.device)
            self.copy_out = torch.cuda.Stream(device=self.device)
        self._pinned = _PinnedPool()
        ctx = C.c_void_p()
        _abi.check(self.lib.b200reg_create(device, C.c_void_p(self.stream.cuda_stream), C.byref(ctx)))
        self.ctx = ctx

    # -- memory -------------------------------------------------------------------------------------
    def empty(self, shape, np_dtype):
        with torch.cuda.stream(self.stream):
            return torch.empty(tuple(int(s) for s in shape), dtype=_DT[np.dtype(np_dtype)][1], device=self.device)

    def zeros(self, shape, np_dtype):
        with torch.cuda.stream(self.stream):
            return torch.zeros(tuple(int(s) for s in shape), dtype=_DT[np.dtype(np_dtype)][1], device=self.device)

    def synchronize(self):
        _abi.check(self.lib.b200reg_synchronize(self.ctx))

    # The engine works on its own stream.  Device handles cross the public API in both directions, so the
    # API wrappers order the engine stream after the caller's current stream on entry (wait_caller) and the
    # caller's current stream after the engine stream on exit (release_to_caller): event waits, no host sync.
    def wait_caller(self):
        self.stream.wait_stream(torch.cuda.current_stream(self.device))

    def release_to_caller(self):
        torch.cuda.current_stream(self.device).wait_stream(self.stream)

    def launch_count(self):
        return int(self.lib.b200reg_launch_count(self.ctx))

    def to_device(self, image):
        """Host ``Image`` -> ``DeviceImage`` (vector images are converted AoS -> SoA on the device)."""
        if isinstance(image, DeviceImage):
            self.wait_caller()
            return image
        image = sk.to_native(image)
        host = _as_torch_host(image.array)
        with torch.cuda.stream(self.stream):
            dev = host.to(self.device, non_blocking=True)
        if image.is_vector:
            if image.array.dtype != np.float64 or image.array.shape[3] != 3:
                raise RuntimeError("only 3-component sitkVectorFloat64 fields are supported")
            n = image.GetNumberOfPixels()
            z, y, x = image.array.shape[:3]
            soa = self.empty((3, z, y, x), np.float64)
            _abi.check(self.lib.b200reg_aos_to_soa(self.ctx, C.c_void_p(dev.data_ptr()), C.c_void_p(soa.data_ptr()), n))
            dev.record_stream(self.stream)
            dev = soa
        return DeviceImage(dev, image.array.dtype, image.GetSpacing(), image.GetOrigin(), image.GetDirection(), image.is_vector)

    def to_device_async(self, image):
        """Host scalar ``Image`` (ideally in pinned memory, see ``pinned_image``) -> ``(DeviceImage, event)``, copied on the copy-in
        stream.  Nothing waits: the consumer calls ``engine.stream.wait_event(event)`` right before the first use, so an upload
        requested early runs beside whatever the engine stream is doing (``event`` is None when there was nothing to copy)."""
        if isinstance(image, DeviceImage):
            return image, None
        image = sk.to_native(image)
        if image.is_vector:
            return self.to_device(image), None
        host = _as_torch_host(image.array)
        with torch.cuda.stream(self.copy_in):
            dev = host.to(self.device, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.copy_in)
        dev.record_stream(self.stream)
        return DeviceImage(dev, image.array.dtype, image.GetSpacing(), image.GetOrigin(), image.GetDirection(), False), ev

    def to_host_async(self, dimg):
        """``DeviceImage`` -> ``(host Image in pinned memory, event)``: the SoA -> AoS conversion of a field runs on the engine stream,
        the PCIe copy on the copy-out stream behind it; nothing waits.  The image is valid once ``event.synchronize()`` returns."""
        x, y, z = dimg.GetSize()
        if dimg.is_vector:
            aos = self.empty((z, y, x, 3), np.float64)
            _abi.check(self.lib.b200reg_soa_to_aos(self.ctx, dimg.ptr, C.c_void_p(aos.data_ptr()), dimg.GetNumberOfPixels()))
            src, shape = aos, (z, y, x, 3)
        else:
            src, shape = dimg.tensor, (z, y, x)
        pinned = self._pinned.take(shape, _DT[dimg.np_dtype][1])
        base = pinned.numpy()  # every view the caller can make keeps this array alive; when it dies the buffer is recycled
        weakref.finalize(base, self._pinned.give, pinned)
        host = base.view(dimg.np_dtype) if dimg.np_dtype in _SIGNED_VIEW else base
        self.copy_out.wait_stream(self.stream)
        src.record_stream(self.copy_out)
        with torch.cuda.stream(self.copy_out):
            pinned.copy_(src.view(pinned.shape), non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.copy_out)
        out = Image.__new__(Image)
        out._arr = host
        out._spacing, out._origin, out._direction, out._is_vector = dimg.spacing, dimg.origin, dimg.direction, dimg.is_vector
        return out, ev

    def to_host(self, dimg, pinned=True):
        """``DeviceImage`` -> host ``Image`` (vector fields come back as AoS ``[z, y, x, 3]``).  Synchronises."""
        x, y, z = dimg.GetSize()
        if dimg.is_vector:
            aos = self.empty((z, y, x, 3), np.float64)
            _abi.check(self.lib.b200reg_soa_to_aos(self.ctx, dimg.ptr, C.c_void_p(aos.data_ptr()), dimg.GetNumberOfPixels()))
            src, shape = aos, (z, y, x, 3)
        else:
            src, shape = dimg.tensor, (z, y, x)
        host = pinned_empty(shape, dimg.np_dtype) if pinned else np.empty(shape, dimg.np_dtype)
        _abi.check(self.lib.b200reg_memcpy_d2h(self.ctx, host.ctypes.data_as(C.c_void_p), C.c_void_p(src.data_ptr()), host.nbytes))
        self.synchronize()
        out = Image.__new__(Image)
        out._arr = host
        out._spacing, out._origin, out._direction, out._is_vector = dimg.spacing, dimg.origin, dimg.direction, dimg.is_vector
        return out

    # -- transform chains ---------------------------------------------------------------------------
    def chain(self, transform):
        """sitk-style transform object -> (ctypes Transform array, n, keep-alive list).  Displacement fields
        are uploaded once and cached on the transform object (reference multiatlas/run.py:331-345 re-uses
        one transform for the CT and every structure)."""
        flat = [] if transform is None else transform.flatten()
        if len(flat) > _abi.MAX_TRANSFORMS:
            raise ValueError(f"transform chain longer than {_abi.MAX_TRANSFORMS}")
        arr = (_abi.Transform * max(len(flat), 1))()
        keep = []
        for i, t in enumerate(flat):
            if isinstance(t, sk.DisplacementFieldTransform):
                cache = t._device_cache
                if cache is None or cache[0] is not self:
                    d = self.to_device(t.GetDisplacementField())
                    t._device_cache = cache = (self, d)
                d = cache[1]
                arr[i].kind = _abi.TFM_DVF
                arr[i].d_dvf = d.tensor.data_ptr()
                arr[i].dvf_geom = d.geom
                keep.append(d)
            elif isinstance(t, sk.AffineTransform):
                arr[i].kind = _abi.TFM_AFFINE
                m, o = t.matrix.reshape(9), t.offset
                for k in range(9):
                    arr[i].matrix[k] = float(m[k])
                for k in range(3):
                    arr[i].offset[k] = float(o[k])
            else:
                raise NotImplementedError(f"transform type {type(t).__name__} is not supported")
        return arr, len(flat), keep

    # -- ops ------------------------------------------------------------------------------------------
    def cast(self, dimg, np_dtype):
        np_dtype = np.dtype(np_dtype)
        if np_dtype == dimg.np_dtype:
            return dimg
        out = self.empty(dimg.tensor.shape, np_dtype)
        _abi.check(self.lib.b200reg_cast(self.ctx, dimg.ptr, dimg.dtype_id, C.c_void_p(out.data_ptr()), _DT[np_dtype][0], dimg.tensor.numel()))
        return dimg.like(out, np_dtype)

    def minmax(self, dimg):
        mn, mx = C.c_double(), C.c_double()
        _abi.check(self.lib.b200reg_minmax(self.ctx, dimg.ptr, dimg.dtype_id, dimg.tensor.numel(), C.byref(mn), C.byref(mx)))
        return mn.value, mx.value

    def discrete_gaussian(self, dimg, variance, maximum_kernel_width=32, maximum_error=0.01, use_image_spacing=True):
        if dimg.np_dtype != np.float32:
            raise NotImplementedError("DiscreteGaussian is implemented for Float32 images")
        var = (C.c_double * 3)(*([float(variance)] * 3 if np.isscalar(variance) else [float(v) for v in variance]))
        out = self.empty(dimg.tensor.shape, np.float32)
        g = dimg.geom
        _abi.check(self.lib.b200reg_discrete_gaussian_f32(self.ctx, dimg.ptr, C.c_void_p(out.data_ptr()), C.byref(g), var,
                                                          int(maximum_kernel_width), float(maximum_error), int(bool(use_image_spacing))))
        return dimg.like(out)

    def smooth_and_resample(self, dimg, variance, maximum_kernel_width, out_geom_src, interpolator=sk.sitkLinear, allow_restricted=True):
        """DiscreteGaussian + Resample onto ``out_geom_src`` of a Float32 image as one library call (utils.py:216-267).  ``allow_restricted``:
        True / 1 = blur only what the resampler reads when the level shrinks enough to pay, 2 = whenever possible, False / 0 = never."""
        if dimg.np_dtype != np.float32 or dimg.is_vector:
            raise NotImplementedError("smooth_and_resample as one call is implemented for Float32 scalar images")
        var = (C.c_double * 3)(*([float(variance)] * 3 if np.isscalar(variance) else [float(v) for v in variance]))
        gin, gout = dimg.geom, _abi.geom_of(out_geom_src)
        x, y, z = out_geom_src.GetSize()
        out = self.empty((z, y, x), np.float32)
        _abi.check(self.lib.b200reg_smooth_and_resample_f32(self.ctx, dimg.ptr, C.byref(gin), var, int(maximum_kernel_width), C.byref(gout),
                                                            int(interpolator), C.c_void_p(out.data_ptr()), int(allow_restricted)))
        return DeviceImage(out, np.float32, out_geom_src.GetSpacing(), out_geom_src.GetOrigin(), out_geom_src.GetDirection(), False)

    def resample_batch(self, images, out_geom_src, transform, interpolators, default_values):
        """N images on one grid through one transform chain onto the grid of ``out_geom_src`` (anything
        with GetSize/GetSpacing/GetOrigin/GetDirection)."""
        n = len(images)
        gin = images[0].geom
        for im in images[1:]:
            if im.GetSize() != images[0].GetSize():
                raise ValueError("resample_batch: all inputs must share one grid")
        gout = _abi.geom_of(out_geom_src)
        x, y, z = out_geom_src.GetSize()
        outs = [self.empty((z, y, x), im.np_dtype) for im in images]
        ins_p = (C.c_void_p * n)(*[im.tensor.data_ptr() for im in images])
        outs_p = (C.c_void_p * n)(*[o.data_ptr() for o in outs])
        dts = (C.c_int * n)(*[im.dtype_id for im in images])
        ips = (C.c_int * n)(*[int(i) for i in interpolators])
        dvs = (C.c_double * n)(*[float(d) for d in default_values])
        chain, nch, keep = self.chain(transform)
        _abi.check(self.lib.b200reg_resample_batch(self.ctx, n, ins_p, dts, C.byref(gin), outs_p, C.byref(gout), chain, nch, ips, dvs))
        sp, og, dr = out_geom_src.GetSpacing(), out_geom_src.GetOrigin(), out_geom_src.GetDirection()
        return [DeviceImage(o, im.np_dtype, sp, og, dr, False) for o, im in zip(outs, images)]

    def resample(self, dimg, out_geom_src=None, transform=None, interpolator=sk.sitkLinear, default_value=0.0):
        if out_geom_src is None:
            out_geom_src = dimg
        if dimg.is_vector:
            return self.resample_vec3(dimg, out_geom_src, transform, default_value)
        return self.resample_batch([dimg], out_geom_src, transform, [interpolator], [default_value])[0]

    def resample_vec3(self, dfield, out_geom_src, transform=None, default_value=0.0):
        gin, gout = dfield.geom, _abi.geom_of(out_geom_src)
        x, y, z = out_geom_src.GetSize()
        out = self.empty((3, z, y, x), np.float64)
        chain, nch, keep = self.chain(transform)
        _abi.check(self.lib.b200reg_resample_vec3(self.ctx, dfield.ptr, C.byref(gin), C.c_void_p(out.data_ptr()), C.byref(gout), chain, nch,
                                                  float(default_value)))
        return DeviceImage(out, np.float64, out_geom_src.GetSpacing(), out_geom_src.GetOrigin(), out_geom_src.GetDirection(), True)

    def transform_to_dvf(self, transform, grid):
        """sitk.TransformToDisplacementField on the grid of ``grid`` (deformable.py:101-108)."""
        x, y, z = grid.GetSize()
        out = self.empty((3, z, y, x), np.float64)
        g = _abi.geom_of(grid)
        chain, nch, keep = self.chain(transform)
        _abi.check(self.lib.b200reg_transform_to_dvf(self.ctx, C.byref(g), chain, nch, C.c_void_p(out.data_ptr())))
        return DeviceImage(out, np.float64, grid.GetSpacing(), grid.GetOrigin(), grid.GetDirection(), True)

    def compose_dvf(self, total, iter_field):
        """total <- total + Resample(iter_field, DisplacementFieldTransform(total))  (deformable.py:154)."""
        scratch = self.empty(total.tensor.shape, np.float64)
        g = total.geom
        _abi.check(self.lib.b200reg_compose_dvf(self.ctx, total.ptr, iter_field.ptr, C.byref(g), C.c_void_p(scratch.data_ptr())))
        return total

    def recursive_gaussian(self, dfield, sigma):
        s = (C.c_double * 3)(*[float(v) for v in sigma])
        g = dfield.geom
        _abi.check(self.lib.b200reg_recursive_gaussian_vec3(self.ctx, dfield.ptr, C.byref(g), s))
        return dfield

    def pde_smooth_field(self, dfield, std_dev, max_error=0.1, max_kernel_width=30):
        s = (C.c_double * 3)(*[float(v) for v in std_dev])
        g = dfield.geom
        _abi.check(self.lib.b200reg_pde_smooth_field(self.ctx, dfield.ptr, C.byref(g), s, float(max_error), int(max_kernel_width)))
        return dfield

    def pyramid_geom(self, src, isotropic, resolution):
        gin, gout = _abi.geom_of(src), _abi.Geom()
        _abi.check(self.lib.b200reg_pyramid_geom(C.byref(gin), int(bool(isotropic)), float(resolution), C.byref(gout)))
        return gout

    def demons_execute(self, fixed, moving, params):
        x, y, z = fixed.GetSize()
        out = self.empty((3, z, y, x), np.float64)
        st = _abi.DemonsStats()
        gf, gm = fixed.geom, moving.geom
        _abi.check(self.lib.b200reg_demons_execute(self.ctx, fixed.ptr, C.byref(gf), moving.ptr, C.byref(gm), C.byref(params),
                                                   C.c_void_p(out.data_ptr()), C.byref(st)))
        stats = {"elapsed_iterations": st.elapsed_iterations, "metric": st.metric, "rms_change": st.rms_change, "gpu_ms": st.gpu_ms,
                 "voxels": fixed.GetNumberOfPixels(), "trace": self.demons_trace(0, st.elapsed_iterations)}
        return fixed.like(out, np.float64, True), stats

    def demons_trace(self, level, n_iterations):
        """[(metric, RMS change)] per iteration of ``level`` of the most recent Demons call: what the reference's IterationEvent
        callbacks see through GetMetric() / GetRMSChange() (deformable.py:260-264, utils.py:37-41)."""
        n = max(int(n_iterations), 0)
        buf = (C.c_double * (2 * max(n, 1)))()
        got = self.lib.b200reg_demons_trace(self.ctx, int(level), buf, n)
        got = min(got, n)
        return [(buf[2 * i], buf[2 * i + 1]) for i in range(got)]

    def demons_force(self, fixed, moving, field, params):
        x, y, z = fixed.GetSize()
        w = self.empty((z, y, x), np.float32)
        u = self.empty((3, z, y, x), np.float64)
        metric, rms = C.c_double(), C.c_double()
        gf, gm = fixed.geom, moving.geom
        _abi.check(self.lib.b200reg_demons_force(self.ctx, fixed.ptr, C.byref(gf), moving.ptr, C.byref(gm), field.ptr, C.byref(params),
                                                 C.c_void_p(w.data_ptr()), C.c_void_p(u.data_ptr()), C.byref(metric), C.byref(rms)))
        return fixed.like(w, np.float32, False), fixed.like(u, np.float64, True), metric.value, rms.value

    def multiscale_demons(self, fixed, moving, cfg, initial_field=None, initial_on_fixed_grid=False):
        x, y, z = fixed.GetSize()
        out = self.empty((3, z, y, x), np.float64)
        stats = (_abi.DemonsStats * max(cfg.n_levels, 1))()
        gf, gm = fixed.geom, moving.geom
        gi = initial_field.geom if (initial_field is not None and not initial_on_fixed_grid) else None
        _abi.check(self.lib.b200reg_multiscale_demons(
            self.ctx, fixed.ptr, C.byref(gf), moving.ptr, C.byref(gm), C.byref(cfg),
            initial_field.ptr if initial_field is not None else None, C.byref(gi) if gi is not None else None,
            C.c_void_p(out.data_ptr()), stats))
        level_stats = [{"elapsed_iterations": s.elapsed_iterations, "metric": s.metric, "rms_change": s.rms_change, "gpu_ms": s.gpu_ms,
                        "voxels": s.voxels_lo, "trace": self.demons_trace(l, s.elapsed_iterations)} for l, s in enumerate(stats[: cfg.n_levels])]
        return fixed.like(out, np.float64, True), level_stats

    # -- fusion ------------------------------------------------------------------------------------------
    def weight_map(self, target, moving, vote_type, factor=1e12, sigma=2.0, epsilon=1e-5):
        out = self.empty(target.tensor.shape, np.float32)
        g = target.geom
        _abi.check(self.lib.b200reg_weight_map(self.ctx, target.ptr, moving.ptr, C.byref(g), int(vote_type), float(factor), float(sigma),
                                               float(epsilon), C.c_void_p(out.data_ptr())))
        return target.like(out, np.float32, False)

    def weight_map_block(self, target, moving, radius, factor, gain):
        out = self.empty(target.tensor.shape, np.float32)
        g = target.geom
        r = (C.c_int32 * 3)(*[int(v) for v in radius])
        _abi.check(self.lib.b200reg_weight_map_block(self.ctx, target.ptr, moving.ptr, C.byref(g), r, float(factor), float(gain),
                                                     C.c_void_p(out.data_ptr())))
        return target.like(out, np.float32, False)

    def normalise_by_max(self, weight, mask=None):
        _abi.check(self.lib.b200reg_normalise_by_max(self.ctx, weight.ptr, mask.ptr if mask is not None else None, weight.tensor.numel()))
        return weight

    def vote_accumulate(self, label, weight, num, den, first):
        _abi.check(self.lib.b200reg_vote_accumulate(self.ctx, label.ptr, weight.ptr, C.c_void_p(num.data_ptr()),
                                                    C.c_void_p(den.data_ptr()) if den is not None else None, label.tensor.numel(), int(bool(first))))

    def vote_finalize(self, num, den, geom_src, smooth_variance, threshold):
        out = self.empty(num.shape, np.float32)
        g = _abi.geom_of(geom_src)
        _abi.check(self.lib.b200reg_vote_finalize(self.ctx, C.c_void_p(num.data_ptr()), C.c_void_p(den.data_ptr()) if den is not None else None,
                                                  C.byref(g), float(smooth_variance), float(threshold or 0.0), C.c_void_p(out.data_ptr())))
        return DeviceImage(out, np.float32, geom_src.GetSpacing(), geom_src.GetOrigin(), geom_src.GetDirection(), False)

    def binary_threshold(self, dimg, lower, upper=255.0):
        out = self.empty(dimg.tensor.shape, np.uint8)
        _abi.check(self.lib.b200reg_binary_threshold(self.ctx, dimg.ptr, dimg.dtype_id, dimg.tensor.numel(), float(lower), float(upper),
                                                     C.c_void_p(out.data_ptr())))
        return dimg.like(out, np.uint8, False)

    def linreg_meansq(self, fixed, moving, total_matrix, total_offset, initial_matrix, center, fixed_mask=None, moving_mask=None, stride=1):
        """Mean-squares accumulators of linear_registration (14 doubles, see include/b200reg.h).  Synchronises."""
        out = (C.c_double * 14)()
        gf, gm = fixed.geom, moving.geom
        d9, d3 = C.c_double * 9, C.c_double * 3
        _abi.check(self.lib.b200reg_linreg_meansq(
            self.ctx, fixed.ptr, C.byref(gf), moving.ptr, C.byref(gm), d9(*np.asarray(total_matrix, float).reshape(9)),
            d3(*np.asarray(total_offset, float).reshape(3)), d9(*np.asarray(initial_matrix, float).reshape(9)),
            d3(*np.asarray(center, float).reshape(3)), fixed_mask.ptr if fixed_mask is not None else None,
            moving_mask.ptr if moving_mask is not None else None, int(stride), out))
        return np.array(out[:], dtype=np.float64)

    def linreg_correlation(self, fixed, moving, total_matrix, total_offset, initial_matrix, center, fixed_mask=None, moving_mask=None, stride=1):
        """The 42 sums behind the correlation metric of linear_registration (see include/b200reg.h).  Synchronises."""
        out = (C.c_double * 42)()
        gf, gm = fixed.geom, moving.geom
        d9, d3 = C.c_double * 9, C.c_double * 3
        _abi.check(self.lib.b200reg_linreg_correlation(
            self.ctx, fixed.ptr, C.byref(gf), moving.ptr, C.byref(gm), d9(*np.asarray(total_matrix, float).reshape(9)),
            d3(*np.asarray(total_offset, float).reshape(3)), d9(*np.asarray(initial_matrix, float).reshape(9)),
            d3(*np.asarray(center, float).reshape(3)), fixed_mask.ptr if fixed_mask is not None else None,
            moving_mask.ptr if moving_mask is not None else None, int(stride), out))
        return np.array(out[:], dtype=np.float64)

    def linreg_mattes_histogram(self, fixed, moving, total_matrix, total_offset, fixed_bins, moving_bins, n_bins=50, fixed_mask=None, moving_mask=None,
                                stride=1):
        """Joint Parzen histogram [n_bins, n_bins] (fixed bin, moving bin) and the number of valid samples.  Synchronises."""
        hist = np.empty((n_bins, n_bins), dtype=np.float64)
        count = C.c_double()
        gf, gm = fixed.geom, moving.geom
        d9, d3, d2 = C.c_double * 9, C.c_double * 3, C.c_double * 2
        _abi.check(self.lib.b200reg_linreg_mattes_histogram(
            self.ctx, fixed.ptr, C.byref(gf), moving.ptr, C.byref(gm), d9(*np.asarray(total_matrix, float).reshape(9)),
            d3(*np.asarray(total_offset, float).reshape(3)), fixed_mask.ptr if fixed_mask is not None else None,
            moving_mask.ptr if moving_mask is not None else None, int(stride), int(n_bins), d2(*[float(v) for v in fixed_bins]),
            d2(*[float(v) for v in moving_bins]), hist.ctypes.data_as(C.POINTER(C.c_double)), C.byref(count)))
        return hist, count.value

    def linreg_mattes_derivative(self, fixed, moving, total_matrix, total_offset, initial_matrix, center, fixed_bins, moving_bins, table, fixed_mask=None,
                                 moving_mask=None, stride=1):
        """[s (3), S (9)] of the Mattes derivative for the host's table log(p / p_M).  Synchronises."""
        table = np.ascontiguousarray(table, dtype=np.float64)
        out = (C.c_double * 12)()
        gf, gm = fixed.geom, moving.geom
        d9, d3, d2 = C.c_double * 9, C.c_double * 3, C.c_double * 2
        _abi.check(self.lib.b200reg_linreg_mattes_derivative(
            self.ctx, fixed.ptr, C.byref(gf), moving.ptr, C.byref(gm), d9(*np.asarray(total_matrix, float).reshape(9)),
            d3(*np.asarray(total_offset, float).reshape(3)), d9(*np.asarray(initial_matrix, float).reshape(9)),
            d3(*np.asarray(center, float).reshape(3)), fixed_mask.ptr if fixed_mask is not None else None,
            moving_mask.ptr if moving_mask is not None else None, int(stride), int(table.shape[0]), d2(*[float(v) for v in fixed_bins]),
            d2(*[float(v) for v in moving_bins]), table.ctypes.data_as(C.POINTER(C.c_double)), out))
        return np.array(out[:], dtype=np.float64)

    def image_moments(self, dimg):
        """[sum v, sum v x, sum v y, sum v z] of a Float32 image in physical coordinates (ImageMomentsCalculator).  Synchronises."""
        out = (C.c_double * 4)()
        g = dimg.geom
        _abi.check(self.lib.b200reg_image_moments(self.ctx, dimg.ptr, C.byref(g), out))
        return np.array(out[:], dtype=np.float64)

    def _size3(self, dimg):
        x, y, z = dimg.GetSize()
        return (C.c_int32 * 3)(x, y, z)

    def binary_fillhole(self, dimg, fully_connected=False):
        out = self.empty(dimg.tensor.shape, np.uint8)
        _abi.check(self.lib.b200reg_binary_fillhole(self.ctx, dimg.ptr, self._size3(dimg), int(bool(fully_connected)), C.c_void_p(out.data_ptr())))
        return dimg.like(out, np.uint8, False)

    def largest_component(self, dimg, fully_connected=False, want_info=False):
        out = self.empty(dimg.tensor.shape, np.uint8)
        ncomp, nvox = C.c_int64(), C.c_int64()
        _abi.check(self.lib.b200reg_largest_component(self.ctx, dimg.ptr, self._size3(dimg), int(bool(fully_connected)), C.c_void_p(out.data_ptr()),
                                                      C.byref(ncomp) if want_info else None, C.byref(nvox) if want_info else None))
        res = dimg.like(out, np.uint8, False)
        return (res, {"n_components": ncomp.value, "voxels": nvox.value}) if want_info else res

    def process_probability(self, dimg, threshold=0.5):
        out = self.empty(dimg.tensor.shape, np.uint8)
        _abi.check(self.lib.b200reg_process_probability(self.ctx, dimg.ptr, dimg.dtype_id, self._size3(dimg), float(threshold),
                                                        C.c_void_p(out.data_ptr()), None))
        return dimg.like(out, np.uint8, False)

    # -- label utilities (utils/crop.py, label/utils.py, multiatlas/run.py:387-437) ----------------------------------
    def bounding_box(self, mask):
        bb = (C.c_int32 * 6)()
        _abi.check(self.lib.b200reg_bounding_box(self.ctx, mask.ptr, self._size3(mask), bb))
        return list(bb)

    def region_copy(self, src, src_index, dst, dst_index, region_size):
        i3 = C.c_int32 * 3
        _abi.check(self.lib.b200reg_region_copy(self.ctx, src.ptr, self._size3(src), i3(*[int(v) for v in src_index]), dst.ptr, self._size3(dst),
                                                i3(*[int(v) for v in dst_index]), i3(*[int(v) for v in region_size]), src.dtype_id))
        return dst

    def resolve_overlap(self, labels_ranked):
        n = len(labels_ranked)
        outs = [self.empty(l.tensor.shape, np.uint8) for l in labels_ranked]
        pin = (C.c_void_p * n)(*[l.tensor.data_ptr() for l in labels_ranked])
        pout = (C.c_void_p * n)(*[o.data_ptr() for o in outs])
        _abi.check(self.lib.b200reg_resolve_overlap(self.ctx, pin, pout, n, labels_ranked[0].tensor.numel()))
        return [l.like(o, np.uint8, False) for l, o in zip(labels_ranked, outs)]

    def binary_closing(self, mask, radius, offsets):
        out = self.empty(mask.tensor.shape, np.uint8)
        offs = np.ascontiguousarray(offsets, dtype=np.int32).reshape(-1, 3)
        _abi.check(self.lib.b200reg_binary_closing(self.ctx, mask.ptr, self._size3(mask), (C.c_int32 * 3)(*[int(v) for v in radius]),
                                                   offs.ctypes.data_as(C.POINTER(C.c_int32)), int(offs.shape[0]), C.c_void_p(out.data_ptr())))
        return mask.like(out, np.uint8, False)

    # -- distance maps, contours, binary morphology, masking (registration/utils.py:270-344, label/projection.py) ------
    def signed_maurer_distance_map(self, mask, inside_is_positive=False, squared_distance=False, use_image_spacing=True):
        """sitk.SignedMaurerDistanceMap of a UInt8 mask (background 0) -> Float32."""
        out = self.empty(mask.tensor.shape, np.float32)
        g = mask.geom
        _abi.check(self.lib.b200reg_signed_maurer_distance_map(self.ctx, mask.ptr, C.byref(g), int(bool(inside_is_positive)),
                                                               int(bool(squared_distance)), int(bool(use_image_spacing)), C.c_void_p(out.data_ptr())))
        return mask.like(out, np.float32, False)

    def label_contour(self, mask, fully_connected=False):
        out = self.empty(mask.tensor.shape, np.uint8)
        _abi.check(self.lib.b200reg_label_contour(self.ctx, mask.ptr, self._size3(mask), int(bool(fully_connected)), C.c_void_p(out.data_ptr())))
        return mask.like(out, np.uint8, False)

    def label_contour_slicewise(self, mask):
        """sitk.LabelContour of every axial slice on its own (in-plane face neighbours)."""
        out = self.empty(mask.tensor.shape, np.uint8)
        _abi.check(self.lib.b200reg_label_contour_slicewise(self.ctx, mask.ptr, self._size3(mask), C.c_void_p(out.data_ptr())))
        return mask.like(out, np.uint8, False)

    def _binary_morph(self, fn, mask, offsets, boundary_to_foreground):
        out = self.empty(mask.tensor.shape, np.uint8)
        offs = np.ascontiguousarray(offsets, dtype=np.int32).reshape(-1, 3)
        _abi.check(fn(self.ctx, mask.ptr, self._size3(mask), offs.ctypes.data_as(C.POINTER(C.c_int32)), int(offs.shape[0]),
                      int(bool(boundary_to_foreground)), C.c_void_p(out.data_ptr())))
        return mask.like(out, np.uint8, False)

    def binary_dilate(self, mask, offsets, boundary_to_foreground=False):
        return self._binary_morph(self.lib.b200reg_binary_dilate, mask, offsets, boundary_to_foreground)

    def binary_erode(self, mask, offsets, boundary_to_foreground=True):
        return self._binary_morph(self.lib.b200reg_binary_erode, mask, offsets, boundary_to_foreground)

    def u8_binary_op(self, a, b, op):
        if a.tensor.shape != b.tensor.shape:
            raise RuntimeError("both images must have the same size")  # ITK: "Inputs do not occupy the same physical space!"
        out = self.empty(a.tensor.shape, np.uint8)
        _abi.check(self.lib.b200reg_u8_binary_op(self.ctx, a.ptr, b.ptr, int(op), C.c_void_p(out.data_ptr()), a.tensor.numel()))
        return a.like(out, np.uint8, False)

    def mask_image(self, dimg, mask, outside_value=0.0):
        """sitk.Mask(image, mask): scalar images of any pixel type and SoA displacement fields."""
        out = self.empty(dimg.tensor.shape, dimg.np_dtype)
        planes = 3 if dimg.is_vector else 1
        _abi.check(self.lib.b200reg_mask_image(self.ctx, dimg.ptr, dimg.dtype_id, mask.ptr, mask.tensor.numel(), planes, float(outside_value),
                                               C.c_void_p(out.data_ptr())))
        return dimg.like(out)

    def divide_scalar(self, dimg, divisor):
        out = self.empty(dimg.tensor.shape, dimg.np_dtype)
        _abi.check(self.lib.b200reg_divide_scalar(self.ctx, dimg.ptr, dimg.dtype_id, dimg.tensor.numel(), float(divisor), C.c_void_p(out.data_ptr())))
        return dimg.like(out)

    def constant_field(self, grid, vector, mask=None):
        """VectorFloat64 field on the grid of ``grid``: ``vector`` (dx, dy, dz) inside ``mask`` (everywhere without), 0 outside."""
        x, y, z = grid.GetSize()
        out = self.empty((3, z, y, x), np.float64)
        _abi.check(self.lib.b200reg_constant_field(self.ctx, mask.ptr if mask is not None else None, x * y * z,
                                                   (C.c_double * 3)(*[float(v) for v in vector]), C.c_void_p(out.data_ptr())))
        return DeviceImage(out, np.float64, grid.GetSpacing(), grid.GetOrigin(), grid.GetDirection(), True)

    def radial_bend_field(self, mask, reference_index, axis, scale, clip_axis=-1, clip_keep_upper=True):
        x, y, z = mask.GetSize()
        out = self.empty((3, z, y, x), np.float64)
        _abi.check(self.lib.b200reg_radial_bend_field(self.ctx, mask.ptr, self._size3(mask), (C.c_int32 * 3)(*[int(v) for v in reference_index]),
                                                      (C.c_double * 3)(*[float(v) for v in axis]), float(scale), int(clip_axis),
                                                      int(bool(clip_keep_upper)), C.c_void_p(out.data_ptr())))
        return mask.like(out, np.float64, True)

    def patch_correlation(self, target, moving, window):
        """Pearson correlation over the window (x, y, z voxels) around every voxel -> Float64 (fusion.py:82-124)."""
        out = self.empty(target.tensor.shape, np.float64)
        _abi.check(self.lib.b200reg_patch_correlation(self.ctx, target.ptr, moving.ptr, self._size3(target), (C.c_int32 * 3)(*[int(v) for v in window]),
                                                      C.c_void_p(out.data_ptr())))
        return target.like(out, np.float64, False)

    def scale_shift(self, dimg, mul=1.0, add=0.0, take_abs=False):
        """(|x| if take_abs else x) * mul + add on a Float32 / Float64 image."""
        if dimg.np_dtype not in (np.dtype(np.float32), np.dtype(np.float64)) or dimg.is_vector:
            raise TypeError("image arithmetic on the device is implemented for scalar Float32 / Float64 images")
        out = self.empty(dimg.tensor.shape, dimg.np_dtype)
        _abi.check(self.lib.b200reg_scale_shift(self.ctx, dimg.ptr, dimg.dtype_id, dimg.tensor.numel(), int(bool(take_abs)), float(mul), float(add),
                                                C.c_void_p(out.data_ptr())))
        return dimg.like(out)

    def pack_decision(self, label, bit, packed, first):
        _abi.check(self.lib.b200reg_pack_decision(self.ctx, label.ptr, int(bit), C.c_void_p(packed.data_ptr()), label.tensor.numel(), int(bool(first))))

    def unpack_decision(self, packed, bit, like):
        out = self.empty(packed.shape, np.uint8)
        _abi.check(self.lib.b200reg_unpack_decision(self.ctx, C.c_void_p(packed.data_ptr()), int(bit), C.c_void_p(out.data_ptr()), packed.numel()))
        return like.like(out, np.uint8, False)

    # -- compact exchange formats of the sharded fusion (multiatlas.py) -----------------------------------------
    def pack_label(self, label, bit, packed, first):
        """packed |= (label != 0) << bit on a UInt8 / UInt16 / UInt32 decision mask (a torch tensor of that width)."""
        _abi.check(self.lib.b200reg_pack_label(self.ctx, label.ptr, int(bit), C.c_void_p(packed.data_ptr()), _MASK_DT[packed.element_size()],
                                               label.tensor.numel(), int(bool(first))))

    def staple_packed(self, packed, holder_mask, like, confidence_weight=1.0, max_iterations=0xFFFFFFFF, threshold=1e-4, rescale=True, want_info=False):
        """sitk.STAPLE + RescaleIntensity + Threshold from a reduced decision mask; raters = bits of ``holder_mask``."""
        out = self.empty(packed.shape, np.float64)
        n = bin(int(holder_mask)).count("1")
        pq = (C.c_double * (2 * n))()
        elapsed = C.c_int32()
        _abi.check(self.lib.b200reg_staple_packed(self.ctx, C.c_void_p(packed.data_ptr()), _MASK_DT[packed.element_size()], C.c_uint32(int(holder_mask)),
                                                  packed.numel(), float(confidence_weight), C.c_uint32(max_iterations), float(threshold or 0.0),
                                                  int(bool(rescale)), C.c_void_p(out.data_ptr()), pq if want_info else None,
                                                  C.byref(elapsed) if want_info else None))
        res = like.like(out, np.float64, False)
        if want_info:
            return res, {"p": list(pq[:n]), "q": list(pq[n:]), "elapsed_iterations": elapsed.value}
        return res

    def count_accumulate(self, label, counts, first, flag):
        """counts (UInt8 tensor) += label values; ``flag`` (int32 tensor, one element) is raised by a label value above 1."""
        _abi.check(self.lib.b200reg_count_accumulate(self.ctx, label.ptr, C.c_void_p(counts.data_ptr()), label.tensor.numel(), int(bool(first)),
                                                     C.c_void_p(flag.data_ptr())))

    def vote_finalize_counts(self, counts, n_holders, geom_src, smooth_variance, threshold):
        out = self.empty(counts.shape, np.float32)
        g = _abi.geom_of(geom_src)
        _abi.check(self.lib.b200reg_vote_finalize_counts(self.ctx, C.c_void_p(counts.data_ptr()), int(n_holders), C.byref(g), float(smooth_variance),
                                                         float(threshold or 0.0), C.c_void_p(out.data_ptr())))
        return DeviceImage(out, np.float32, geom_src.GetSpacing(), geom_src.GetOrigin(), geom_src.GetDirection(), False)

    def staple(self, decisions, confidence_weight=1.0, max_iterations=0xFFFFFFFF, threshold=1e-4, rescale=True):
        n = len(decisions)
        ptrs = (C.c_void_p * n)(*[d.tensor.data_ptr() for d in decisions])
        out = self.empty(decisions[0].tensor.shape, np.float64)
        pq = (C.c_double * (2 * n))()
        elapsed = C.c_int32()
        _abi.check(self.lib.b200reg_staple(self.ctx, ptrs, n, decisions[0].tensor.numel(), float(confidence_weight), C.c_uint32(max_iterations),
                                           float(threshold or 0.0), int(bool(rescale)), C.c_void_p(out.data_ptr()), pq, C.byref(elapsed)))
        info = {"p": list(pq[:n]), "q": list(pq[n:]), "elapsed_iterations": elapsed.value}
        return decisions[0].like(out, np.float64, False), info
