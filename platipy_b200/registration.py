"""
Drop-in replacements for the reference's registration entry points, running on one B200.

    fast_symmetric_forces_demons_registration   platipy/imaging/registration/deformable.py:190-306
    multiscale_demons                           platipy/imaging/registration/deformable.py:31-187
    apply_transform / apply_linear_transform / apply_deformable_transform
                                                platipy/imaging/registration/utils.py:54-192
    smooth_and_resample                         platipy/imaging/registration/utils.py:195-267
    FastSymmetricForcesDemonsRegistrationFilter the method set of the SimpleITK filter the reference uses
                                                (deformable.py:244-257,143-149,157; utils.py:41)

Signatures, defaults, argument meaning and error classes follow the reference.  Inputs may be host images
(``sitk_compat.Image`` or real ``SimpleITK.Image``) or ``DeviceImage`` handles; host inputs give host outputs
of the same kind, device inputs give device outputs (nothing leaves HBM).
"""
from __future__ import annotations

import logging

import numpy as np

from . import _abi
from . import sitk_compat as sk
from .engine import DeviceImage, Engine

logger = logging.getLogger(__name__)

sitkNearestNeighbor, sitkLinear, sitkBSpline = sk.sitkNearestNeighbor, sk.sitkLinear, sk.sitkBSpline

# "parity" (Float64 fields, the reference's arithmetic) or "fast" (float32 fields inside the Demons loop); every registration entry
# point also takes ``precision=`` for one call
DEFAULT_PRECISION = "parity"

# per-level statistics of the most recent multiscale_demons call (elapsed iterations, metric, RMS change, device ms)
LAST_LEVEL_STATS = []


def _check_interp(interpolator):
    # SimpleITK enum values (deformable.py:221-224): 1 nearest neighbour, 2 linear, 3 B-spline (order 3)
    if interpolator not in (sk.sitkNearestNeighbor, sk.sitkLinear, sk.sitkBSpline):
        raise NotImplementedError(f"interpolator {interpolator!r} is not implemented on the B200 path "
                                  "(sitkNearestNeighbor, sitkLinear and sitkBSpline are)")
    return int(interpolator)


def _is_device(image):
    return isinstance(image, DeviceImage)


def _back(eng, dimg, like):
    """Return ``dimg`` in the form of ``like``: device handle, stand-in Image or real SimpleITK image."""
    if _is_device(like):
        eng.release_to_caller()
        return dimg
    return sk.from_native(eng.to_host(dimg), like)


# ---------------------------------------------------------------------------------------------------------
# the small helpers of registration/utils.py:22-51 (iteration callbacks, control-point arithmetic)
# ---------------------------------------------------------------------------------------------------------
def registration_command_iteration(method):
    """Print one line per optimiser iteration of a linear registration (utils.py:22-27)."""
    print("{0:3} = {1:10.5f}".format(method.GetOptimizerIteration(), method.GetMetricValue()))


def stage_iteration(method):
    """Print the number of transform parameters at a stage change (utils.py:30-34)."""
    print(f"Number of parameters = {method.GetInitialTransform().GetNumberOfParameters()}")


def deformable_registration_command_iteration(method):
    """Print one line per Demons level / iteration event (utils.py:37-41); works with this module's filter object."""
    print("{0:3} = {1:10.5f}".format(method.GetElapsedIterations(), method.GetMetric()))


def control_point_spacing_distance_to_number(image, grid_spacing):
    """Grid spacing in millimetres -> number of control points per axis (utils.py:44-51)."""
    image_spacing = np.array(image.GetSpacing())
    image_size = np.array(image.GetSize())
    number_points = image_size * image_spacing / np.array(grid_spacing)
    return (number_points + 0.5).astype(int)


# ---------------------------------------------------------------------------------------------------------
# smooth_and_resample (utils.py:195-267)
# ---------------------------------------------------------------------------------------------------------
def smooth_and_resample(image, isotropic_voxel_size_mm=None, shrink_factor=None, smoothing_sigma=None,
                        interpolator=sitkLinear):
    eng = Engine.get()
    d = eng.to_device(image)
    pending_blur = None
    if smoothing_sigma:
        if hasattr(smoothing_sigma, "__iter__"):
            variance = [s * s for s in smoothing_sigma]
        else:
            variance = (smoothing_sigma ** 2,) * 3
        max_width = int(max(8 * v * sp for sp, v in zip(d.GetSpacing(), variance)))
        if d.np_dtype == np.float32 and not d.is_vector and (isotropic_voxel_size_mm or shrink_factor):
            # blur + resample go to the library as one call: a shrinking level is blurred only where the resampler reads it (bit-identical)
            pending_blur = (variance, max_width)
        else:
            d = eng.discrete_gaussian(d, variance, max_width)

    size_o, spacing_o = d.GetSize(), d.GetSpacing()
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
        return _back(eng, d, image)
    spacing_n = [((so - 1) * sp) / (sn - 1) for so, sp, sn in zip(size_o, spacing_o, size_n)]

    class _Grid:
        def GetSize(self):
            return tuple(size_n)

        def GetSpacing(self):
            return tuple(spacing_n)

        def GetOrigin(self):
            return d.GetOrigin()

        def GetDirection(self):
            return d.GetDirection()

    if pending_blur is not None:
        out = eng.smooth_and_resample(d, pending_blur[0], pending_blur[1], _Grid(), _check_interp(interpolator))
    else:
        out = eng.resample(d, _Grid(), None, _check_interp(interpolator), 0.0)
    return _back(eng, out, image)


# ---------------------------------------------------------------------------------------------------------
# apply_transform family (utils.py:54-192)
# ---------------------------------------------------------------------------------------------------------
def apply_transform(input_image, reference_image=None, transform=None, default_value=0, interpolator=sitkNearestNeighbor):
    """Resample ``input_image`` onto ``reference_image`` (or its own grid) through ``transform``.
    Output pixel type = input pixel type (utils.py:174,190)."""
    eng = Engine.get()
    d = eng.to_device(input_image)
    ref = reference_image if reference_image else input_image
    out = eng.resample(d, ref, transform, _check_interp(interpolator), default_value)
    return _back(eng, out, input_image)


def apply_transform_batch(input_images, reference_image=None, transform=None, default_values=None, interpolators=None):
    """Batched form of the reference's call pattern (multiatlas/run.py:331-345): one CT plus S label masks
    on the same grid through the same transform; the transform (and the displacement field behind it) is
    evaluated once per output voxel instead of S+1 times."""
    eng = Engine.get()
    n = len(input_images)
    default_values = [0] * n if default_values is None else list(default_values)
    interpolators = [sitkNearestNeighbor] * n if interpolators is None else list(interpolators)
    ds = [eng.to_device(im) for im in input_images]
    ref = reference_image if reference_image else input_images[0]
    outs = []
    for s in range(0, n, _abi.MAX_BATCH):
        outs += eng.resample_batch(ds[s:s + _abi.MAX_BATCH], ref, transform, [_check_interp(i) for i in interpolators[s:s + _abi.MAX_BATCH]],
                                   default_values[s:s + _abi.MAX_BATCH])
    return [_back(eng, o, im) for o, im in zip(outs, input_images)]


def apply_linear_transform(input_image, reference_image, transform, is_structure=False, default_value=0,
                           interpolator=sitkNearestNeighbor):
    if is_structure:
        if default_value != 0 or interpolator != sitkNearestNeighbor:
            logger.warning("is_structure is set to True, but you have set default_value and/or interpolator. "
                           "default_value and/or interpolator will be overwritten.")
        default_value = 0
        interpolator = sitkNearestNeighbor
    return apply_transform(input_image=input_image, reference_image=reference_image, transform=transform,
                           default_value=default_value, interpolator=interpolator)


def apply_deformable_transform(input_image, transform, is_structure=False, default_value=0, interpolator=sitkNearestNeighbor):
    if is_structure:
        if default_value != 0 or interpolator != sitkNearestNeighbor:
            logger.warning("is_structure is set to True, but you have set default_value and/or interpolator. "
                           "default_value and/or interpolator will be overwritten.")
        default_value = 0
        interpolator = sitkNearestNeighbor
    return apply_transform(input_image=input_image, reference_image=None, transform=transform,
                           default_value=default_value, interpolator=interpolator)


# ---------------------------------------------------------------------------------------------------------
# The Demons filter object (duck-types sitk.FastSymmetricForcesDemonsRegistrationFilter)
# ---------------------------------------------------------------------------------------------------------
class FastSymmetricForcesDemonsRegistrationFilter:
    """SimpleITK defaults: StandardDeviations 1.0, NumberOfIterations 10, MaximumRMSError 0.02,
    MaximumUpdateStepLength 0.5, SmoothDisplacementField on, SmoothUpdateField off,
    UpdateFieldStandardDeviations 1.0, MaximumKernelWidth 30, MaximumError 0.1,
    IntensityDifferenceThreshold 0.001.  ``Execute`` can be handed to the reference's own
    ``multiscale_demons`` (it needs SetNumberOfIterations / Execute / GetStandardDeviations)."""

    def __init__(self):
        self._std = [1.0, 1.0, 1.0]
        self._ustd = [1.0, 1.0, 1.0]
        self._iters = 10
        self._smooth_field = True
        self._smooth_update = False
        self._max_rms = 0.02
        self._max_step = 0.5
        self._max_kw = 30
        self._max_err = 0.1
        self._idt = 0.001
        self._precision = 0
        self._commands = []
        self._stats = {"elapsed_iterations": 0, "metric": 0.0, "rms_change": 0.0}

    # setters / getters used by the reference ---------------------------------------------------------
    def SetNumberOfThreads(self, n):  # meaningless on a GPU; accepted and ignored (deformable.py:247)
        pass

    def SetSmoothUpdateField(self, flag):
        self._smooth_update = bool(flag)

    def SetSmoothDisplacementField(self, flag):
        self._smooth_field = bool(flag)

    @staticmethod
    def _vec3(v):
        return [float(v)] * 3 if np.isscalar(v) else [float(s) for s in v]

    def SetStandardDeviations(self, sd):
        self._std = self._vec3(sd)

    def GetStandardDeviations(self):
        return tuple(self._std)

    def SetUpdateFieldStandardDeviations(self, sd):
        self._ustd = self._vec3(sd)

    def SetNumberOfIterations(self, n):
        self._iters = int(n)

    def GetNumberOfIterations(self):
        return self._iters

    def SetMaximumRMSError(self, v):
        self._max_rms = float(v)

    def SetMaximumUpdateStepLength(self, v):
        self._max_step = float(v)

    def SetMaximumKernelWidth(self, v):
        self._max_kw = int(v)

    def SetMaximumError(self, v):
        self._max_err = float(v)

    def SetIntensityDifferenceThreshold(self, v):
        self._idt = float(v)

    def SetFieldPrecision(self, precision):
        """``"parity"`` (default): Float64 fields like ITK's, results comparable bit for bit with the CPU oracle.  ``"fast"``: float32
        fields and float32 FMA smoothing inside the Demons loop -- about half the HBM traffic per iteration, NOT a parity
        path: expect differences of 1e-5 .. 1e-3 mm against the parity result (``bench.py`` reports the percentiles)."""
        if precision not in ("parity", "fast"):
            raise ValueError("precision must be 'parity' or 'fast'")
        self._precision = 1 if precision == "fast" else 0

    def AddCommand(self, event, callback):
        """Callbacks registered for sitkIterationEvent (or sitkAnyEvent) run once per Demons iteration; see _fire_iteration_events."""
        if event in (sk.sitkIterationEvent, sk.sitkAnyEvent):
            self._commands.append(callback)

    def GetElapsedIterations(self):
        return self._stats["elapsed_iterations"]

    def GetMetric(self):
        return self._stats["metric"]

    def GetRMSChange(self):
        return self._stats["rms_change"]

    def params(self, iterations=None):
        p = _abi.DemonsParams()
        for i in range(3):
            p.std_dev[i] = self._std[i]
            p.update_std_dev[i] = self._ustd[i]
        p.smooth_displacement_field = int(self._smooth_field)
        p.smooth_update_field = int(self._smooth_update)
        p.max_error = self._max_err
        p.max_kernel_width = self._max_kw
        p.number_of_iterations = self._iters if iterations is None else int(iterations)
        p.max_rms_error = self._max_rms
        p.max_update_step_length = self._max_step
        p.intensity_difference_threshold = self._idt
        p.denominator_threshold = 1e-9
        p.field_precision = self._precision
        return p

    def Execute(self, fixed_image, moving_image):
        eng = Engine.get()
        f, m = eng.to_device(fixed_image), eng.to_device(moving_image)
        if f.np_dtype != np.float32 or m.np_dtype != np.float32:
            raise RuntimeError("FastSymmetricForcesDemonsRegistrationFilter: fixed and moving images must be sitkFloat32")
        dvf, stats = eng.demons_execute(f, m, self.params())
        self._fire_iteration_events(stats)
        return _back(eng, dvf, fixed_image)

    def _fire_iteration_events(self, stats):
        """The reference's callbacks run at every sitkIterationEvent and read GetElapsedIterations() / GetMetric() there
        (deformable.py:260-264, utils.py:37-41).  The whole loop runs on the device without host round trips, so the events are
        replayed from the recorded per-iteration trace once the level has finished: same sequence of values, later in time."""
        if self._commands:
            for i, (metric, rms) in enumerate(stats.get("trace", [])):
                self._stats = {"elapsed_iterations": i + 1, "metric": metric, "rms_change": rms}
                for cb in self._commands:
                    cb()
        self._stats = stats


B200DemonsFilter = FastSymmetricForcesDemonsRegistrationFilter


def _multires_config(reg, resolution_staging, smoothing_sigmas, iteration_staging, isotropic_resample, interp_order):
    n = len(resolution_staging)
    if not (len(smoothing_sigmas) >= n and len(iteration_staging) >= n):
        raise ValueError("resolution_staging, smoothing_sigmas and iteration_staging must have one entry per level")
    if n > _abi.MAX_LEVELS:
        raise ValueError(f"at most {_abi.MAX_LEVELS} resolution levels are supported")
    cfg = _abi.MultiresConfig()
    cfg.n_levels = n
    cfg.isotropic_resample = int(bool(isotropic_resample))
    for i in range(n):
        cfg.resolution_staging[i] = float(resolution_staging[i] or 0.0)
        cfg.smoothing_sigmas[i] = float(smoothing_sigmas[i] or 0.0)
        cfg.iteration_staging[i] = int(iteration_staging[i])
    cfg.interp_order = _check_interp(interp_order)
    cfg.demons = reg.params()
    return cfg


def multiscale_demons(registration_algorithm, fixed_image, moving_image, initial_transform=None,
                      initial_displacement_field=None, isotropic_resample=None, resolution_staging=None,
                      smoothing_sigmas=None, iteration_staging=None, interp_order=sitkLinear):
    """Multi-resolution driver (deformable.py:31-187).  With the B200 filter the whole pyramid / level loop
    runs inside one C-ABI call and nothing leaves the device; returns the displacement field
    (VectorFloat64 on the fixed grid)."""
    if not isinstance(registration_algorithm, FastSymmetricForcesDemonsRegistrationFilter):
        raise TypeError("multiscale_demons on the B200 path needs platipy_b200's FastSymmetricForcesDemonsRegistrationFilter "
                        "(hand that filter to the reference's multiscale_demons to drive it from SimpleITK code)")
    eng = Engine.get()
    f, m = eng.to_device(fixed_image), eng.to_device(moving_image)
    if f.np_dtype != np.float32 or m.np_dtype != np.float32:
        raise RuntimeError("multiscale_demons: fixed and moving images must be sitkFloat32")
    init, init_on_fixed_grid = None, False
    if initial_displacement_field:
        init = eng.to_device(initial_displacement_field)
    elif initial_transform:
        init, init_on_fixed_grid = eng.transform_to_dvf(initial_transform, f), True  # deformable.py:101-108
    cfg = _multires_config(registration_algorithm, resolution_staging, smoothing_sigmas, iteration_staging, isotropic_resample, interp_order)
    dvf, level_stats = eng.multiscale_demons(f, m, cfg, init, init_on_fixed_grid)
    registration_algorithm.level_stats = level_stats
    LAST_LEVEL_STATS[:] = level_stats
    for st in level_stats:  # IterationEvent callbacks, level by level (deformable.py:143-149 runs the filter once per level)
        registration_algorithm._fire_iteration_events(st)
    return _back(eng, dvf, fixed_image)


def _register_on_device(eng, f0, m0, resolution_staging, iteration_staging, isotropic_resample, initial_displacement_field,
                        regularisation_kernel_mm, smoothing_sigma_factor, smoothing_sigmas, default_value, ncores, interp_order, verbose,
                        on_field_ready=None, precision=None):
    """Device part of fast_symmetric_forces_demons_registration: ``DeviceImage`` inputs -> ``(registered, transform, field)`` on the
    device.  ``on_field_ready(dvf)`` is called as soon as the field is final, before the last warp is enqueued (the host API starts
    the field's PCIe copy there, so that it overlaps the warp)."""
    moving_dtype = m0.np_dtype
    # deformable.py:238-241: cast to Float32 unless the pixel id is 6 (Int64 -- the reference's quirk)
    if f0.GetPixelID() == 6 or m0.GetPixelID() == 6:
        raise RuntimeError("FastSymmetricForcesDemonsRegistrationFilter: Int64 images are not cast by the reference "
                           "(GetPixelID() != 6 test) and the ITK filter then rejects them")
    f = eng.cast(f0, np.float32)
    m = eng.cast(m0, np.float32)

    reg = FastSymmetricForcesDemonsRegistrationFilter()
    reg.SetNumberOfThreads(ncores)
    reg.SetSmoothUpdateField(True)
    reg.SetSmoothDisplacementField(True)
    reg.SetFieldPrecision(precision or DEFAULT_PRECISION)
    # deformable.py:253-257: voxel-unit sigmas from the full-resolution spacing, reused at every level
    reg.SetStandardDeviations((np.array(regularisation_kernel_mm) / np.array(f.GetSpacing())).tolist())

    if not smoothing_sigmas:
        smoothing_sigmas = [i * smoothing_sigma_factor for i in resolution_staging]

    if verbose:
        # deformable.py:260-264: one line per iteration, "{elapsed:3} = {metric:10.5f}"
        reg.AddCommand(sk.sitkIterationEvent, lambda: deformable_registration_command_iteration(reg))

    # deformable.py:286-293: CT-like default value of the final resample (asked for here, ahead of the long device-resident loop,
    # because the answer costs a host synchronisation)
    if default_value is None:
        default_value = 0
        if eng.minmax(m)[0] <= -1000:
            default_value = -1000

    dvf = multiscale_demons(reg, f, m, resolution_staging=resolution_staging, smoothing_sigmas=smoothing_sigmas,
                            iteration_staging=iteration_staging, isotropic_resample=isotropic_resample,
                            initial_displacement_field=initial_displacement_field, interp_order=interp_order)
    if on_field_ready is not None:
        on_field_ready(dvf)

    tfm = sk.DisplacementFieldTransform.__new__(sk.DisplacementFieldTransform)
    tfm._field = None
    tfm._device_cache = (eng, dvf)
    # deformable.py:281-304: final resample on the fixed grid, cast back to the moving image's pixel type
    reg_img = eng.resample(m, f, tfm, _check_interp(interp_order), default_value)
    reg_img = eng.cast(reg_img, moving_dtype)
    return reg_img, tfm, dvf


class PendingRegistration:
    """Handle returned by ``submit_registration``: the registration has run on the device, its results are on their way to pinned
    host memory on the copy-out stream.  ``result()`` waits for the copies and returns what the synchronous call returns."""

    def __init__(self, fixed_like, moving_like, tfm, image_host, image_event, field_host, field_event, level_stats):
        self._like = (fixed_like, moving_like)
        self._tfm, self._image, self._field = tfm, image_host, field_host
        self._events = (image_event, field_event)
        self.level_stats = level_stats

    def done(self):
        return all(e.query() for e in self._events)

    def result(self):
        for e in self._events:
            e.synchronize()
        self._tfm._field = self._field
        return sk.from_native(self._image, self._like[1]), self._tfm, sk.from_native(self._field, self._like[0])


def submit_registration(fixed_image, moving_image, resolution_staging=[8, 4, 1], iteration_staging=[10, 10, 10], isotropic_resample=False,
                        initial_displacement_field=None, regularisation_kernel_mm=1.5, smoothing_sigma_factor=1, smoothing_sigmas=False,
                        default_value=None, ncores=1, interp_order=sitkLinear, verbose=False, _uploaded=None, precision=None):
    """``fast_symmetric_forces_demons_registration`` for host images whose results are not needed at once: the inputs go up on the
    copy-in stream, the registration runs, and the 1.9 GB of results (VectorFloat64 field + registered image at 512 x 512 x 256) come
    down on the copy-out stream while the caller goes on -- typically to the next ``submit_registration``, whose uploads and
    compute then overlap these downloads (the reference's own use is a loop over atlases, multiatlas/run.py:312-347).
    Returns a ``PendingRegistration``."""
    eng = Engine.get()
    (f0, f_ready), (m0, m_ready) = _uploaded if _uploaded is not None else (eng.to_device_async(fixed_image), eng.to_device_async(moving_image))
    for ev in (f_ready, m_ready):
        if ev is not None:
            eng.stream.wait_event(ev)
    started = {}

    def start_field_copy(dvf):
        started["field"] = eng.to_host_async(dvf)

    reg_img, tfm, dvf = _register_on_device(eng, f0, m0, resolution_staging, iteration_staging, isotropic_resample, initial_displacement_field,
                                            regularisation_kernel_mm, smoothing_sigma_factor, smoothing_sigmas, default_value, ncores, interp_order,
                                            verbose, on_field_ready=start_field_copy, precision=precision)
    image_host, image_event = eng.to_host_async(reg_img)
    field_host, field_event = started["field"]
    return PendingRegistration(fixed_image, moving_image, tfm, image_host, image_event, field_host, field_event, LAST_LEVEL_STATS[:])


def iter_registrations(pairs, **kwargs):
    """Registrations of a sequence of ``(fixed, moving)`` host pairs, pipelined over PCIe: while pair k is being registered the
    inputs of pair k + 1 are uploaded and the results of pair k - 1 downloaded.  Yields ``(image, transform, field)`` per pair, in
    order, each as soon as its copies have landed -- the same values as a loop of ``fast_symmetric_forces_demons_registration``
    calls.  At most two results are in flight, so a long sequence needs no more pinned memory than a short one."""
    eng = Engine.get()
    it = iter(pairs)
    cur_pair = next(it, None)
    cur_up = (eng.to_device_async(cur_pair[0]), eng.to_device_async(cur_pair[1])) if cur_pair is not None else None
    pending = None
    while cur_pair is not None:
        nxt_pair = next(it, None)
        # the uploads of the next pair are queued before this pair's compute is: they run beside it
        nxt_up = (eng.to_device_async(nxt_pair[0]), eng.to_device_async(nxt_pair[1])) if nxt_pair is not None else None
        p = submit_registration(cur_pair[0], cur_pair[1], _uploaded=cur_up, **kwargs)
        if pending is not None:
            yield pending.result()
        pending, cur_pair, cur_up = p, nxt_pair, nxt_up
    if pending is not None:
        yield pending.result()


def register_batch(pairs, **kwargs):
    """``list(iter_registrations(pairs, **kwargs))``."""
    return list(iter_registrations(pairs, **kwargs))


def fast_symmetric_forces_demons_registration(fixed_image, moving_image, resolution_staging=[8, 4, 1],
                                              iteration_staging=[10, 10, 10], isotropic_resample=False,
                                              initial_displacement_field=None, regularisation_kernel_mm=1.5,
                                              smoothing_sigma_factor=1, smoothing_sigmas=False, default_value=None,
                                              ncores=1, interp_order=sitkLinear, verbose=False, precision=None):
    """Deformable image propagation using Fast Symmetric-Forces Demons (deformable.py:190-306).  ``precision`` is this package's one
    extra (keyword) argument: ``"parity"`` (default) or ``"fast"``, see ``FastSymmetricForcesDemonsRegistrationFilter.SetFieldPrecision``.

    Returns ``(registered_image, DisplacementFieldTransform, displacement_field)``.  For ``DeviceImage``
    inputs the three results stay on the device (the transform object then wraps the device field).  Host inputs give host
    results; ``submit_registration`` / ``register_batch`` are the forms that overlap the PCIe copies of back-to-back calls.
    """
    eng = Engine.get()
    if _is_device(fixed_image) and _is_device(moving_image):
        f0, m0 = eng.to_device(fixed_image), eng.to_device(moving_image)
        reg_img, tfm, dvf = _register_on_device(eng, f0, m0, resolution_staging, iteration_staging, isotropic_resample, initial_displacement_field,
                                                regularisation_kernel_mm, smoothing_sigma_factor, smoothing_sigmas, default_value, ncores,
                                                interp_order, verbose, precision=precision)
        tfm._field = dvf
        eng.release_to_caller()
        return reg_img, tfm, dvf
    if _is_device(fixed_image) or _is_device(moving_image) or sk.to_native(fixed_image).is_vector:
        # mixed representations: the plain synchronous path
        f0, m0 = eng.to_device(fixed_image), eng.to_device(moving_image)
        reg_img, tfm, dvf = _register_on_device(eng, f0, m0, resolution_staging, iteration_staging, isotropic_resample, initial_displacement_field,
                                                regularisation_kernel_mm, smoothing_sigma_factor, smoothing_sigmas, default_value, ncores,
                                                interp_order, verbose, precision=precision)
        dvf_host = eng.to_host(dvf)
        tfm._field = dvf_host
        return sk.from_native(eng.to_host(reg_img), moving_image), tfm, sk.from_native(dvf_host, fixed_image)
    return submit_registration(fixed_image, moving_image, resolution_staging, iteration_staging, isotropic_resample, initial_displacement_field,
                               regularisation_kernel_mm, smoothing_sigma_factor, smoothing_sigmas, default_value, ncores, interp_order,
                               verbose, precision=precision).result()


# ---------------------------------------------------------------------------------------------------------
# field exponentiation (scaling and squaring)
# ---------------------------------------------------------------------------------------------------------
def exponentiate_field(velocity_field, number_of_iterations=None, maximum_number_of_iterations=20):
    """exp(v) of a stationary velocity field by scaling and squaring, the scheme of ITK's ExponentialDisplacementFieldImageFilter:
    with N squarings, u_0 = v / 2^N and u_{k+1}(x) = u_k(x) + u_k(x + u_k(x)) (linear interpolation, identity outside the grid).
    ``number_of_iterations=None`` picks N like the filter's automatic mode: the smallest N >= 0 with max|v| / 2^N <= half the
    smallest voxel spacing, capped at ``maximum_number_of_iterations``.

    The reference path itself never exponentiates: FastSymmetricForcesDemons is the additive ESM scheme and platipy only *composes*
    level fields (deformable.py:154).  This is the optional diffeomorphic-update building block BASELINE.json's
    north star names, built from the same composition kernel; it is not on any parity path."""
    import torch  # plumbing: one max-norm reduction to choose N

    eng = Engine.get()
    v = eng.to_device(velocity_field)
    if not v.is_vector:
        raise RuntimeError("exponentiate_field expects a sitkVectorFloat64 field")
    if number_of_iterations is None:
        with torch.cuda.stream(eng.stream):
            max_norm = float(torch.sqrt((v.tensor * v.tensor).sum(dim=0)).max().item())
        half_voxel = 0.5 * min(v.GetSpacing())
        n = 0
        while max_norm / (2.0 ** n) > half_voxel and n < int(maximum_number_of_iterations):
            n += 1
    else:
        n = int(number_of_iterations)
    if n > 0:
        u = eng.divide_scalar(v, 2.0 ** n)
    else:
        with torch.cuda.stream(eng.stream):
            u = v.like(v.tensor.clone())
    for _ in range(n):
        eng.compose_dvf(u, u)  # u <- u + Resample(u, DisplacementFieldTransform(u)): one resample_vec3 with the accumulate form
    return _back(eng, u, velocity_field)

