#!/usr/bin/env python
"""
validate_against_sitk.py -- pin the CPU oracle (and, with --gpu, the CUDA path) against the REAL reference.

The arithmetic of platipy's hot path lives in SimpleITK 2.3.1 / ITK 5.3, which is not installable in the build container, so
`oracle/` restates it from memory of the ITK sources ("parity unpinned").  This script is what turns
"unpinned" into "pinned" the moment an environment with SimpleITK exists:

    pip install SimpleITK==2.3.1        # the version platipy locks (poetry.lock:4523-4524)
    python tools/validate_against_sitk.py [--reference /path/to/platipy/checkout] [--gpu] [--json report.json]

It imports the reference's OWN modules -- platipy/imaging/registration/deformable.py:190, registration/utils.py:148,195,
label/fusion.py:56,205,239,295 -- from the checkout (the unused plotting import of deformable.py:27 and the scikit-image import
of fusion.py:21 are stubbed when those packages are missing), runs them on the synthetic inputs of SURVEY.md section 8d, and
compares every result with the oracle's restatement of the same call:

  * per ITK filter: GaussianOperator coefficients, DiscreteGaussian, Resample (identity / affine / displacement field, nearest
    neighbour / linear, scalar and vector images), DisplacementFieldTransform, SmoothingRecursiveGaussian, BinaryThreshold,
    STAPLE, the FastSymmetricForcesDemons filter itself (elapsed iterations, metric, field);
  * per platipy function: smooth_and_resample, apply_transform, fast_symmetric_forces_demons_registration (cfg1 and a 3-level
    case), compute_weight_map, combine_labels, combine_labels_staple, process_probability_image.

For every named semantic switch (oracle/itk_oracle.c g_semantics == include/b200reg.h b200reg_set_semantic) the governed check
is run under BOTH settings and the report says which one matches SimpleITK: a mismatch of a recalled default is then fixed by
flipping the switch on both sides, not by editing CUDA code.  Exit status: 0 all checks pass with the default switches,
1 some check fails, 3 SimpleITK is not importable (nothing was validated).

TEST INFRASTRUCTURE: nothing in platipy_b200/ imports this file.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

DVF_TOL_MM = 1e-4   # north star: DVF within 1e-4 mm per component
REL_TOL = 1e-5      # resampled float intensities within 1e-5 relative


def have_simpleitk():
    try:
        import SimpleITK  # noqa: F401

        return True
    except Exception:  # noqa: BLE001
        return False


def import_reference(reference_root="/root/reference"):
    """Import the reference's hot-path modules from a checkout.  Only what the hot path does not use is stubbed:
    ``platipy.imaging.visualisation.visualiser`` (matplotlib; deformable.py:27 imports ImageVisualiser for commented-out code,
    deformable.py:161-183) and ``skimage.util.shape`` (fusion.py:21, used by the patch-correlation vote only) when missing.
    Returns a namespace with the reference functions."""
    if reference_root not in sys.path:
        sys.path.insert(0, reference_root)
    try:
        import matplotlib  # noqa: F401
    except Exception:  # noqa: BLE001
        # plotting is never reached on the hot path: deformable.py:27 imports ImageVisualiser for commented-out code, label/utils.py:20
        # pulls utils/math.py, whose one matplotlib use is an optional figure
        class _Plot(types.ModuleType):
            __path__ = []

            def __getattr__(self, name):
                if name.startswith("__"):
                    raise AttributeError(name)

                def _unused(*a, **k):
                    raise RuntimeError("plotting is stubbed in validate_against_sitk.py")

                return _unused

        for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.colors", "matplotlib.cm", "matplotlib.patches", "matplotlib.lines",
                     "matplotlib.ticker", "matplotlib.gridspec"):
            sys.modules.setdefault(name, _Plot(name))
        pkg = types.ModuleType("platipy.imaging.visualisation")
        pkg.__path__ = []
        vis = types.ModuleType("platipy.imaging.visualisation.visualiser")

        class ImageVisualiser:  # never instantiated on the hot path
            def __init__(self, *a, **k):
                raise RuntimeError("plotting is stubbed in validate_against_sitk.py")

        vis.ImageVisualiser = ImageVisualiser
        sys.modules.setdefault("platipy.imaging.visualisation", pkg)
        sys.modules.setdefault("platipy.imaging.visualisation.visualiser", vis)
    try:
        import skimage.util.shape  # noqa: F401
    except Exception:  # noqa: BLE001
        sk_pkg, sk_util, sk_shape = types.ModuleType("skimage"), types.ModuleType("skimage.util"), types.ModuleType("skimage.util.shape")
        sk_pkg.__path__, sk_util.__path__ = [], []

        def view_as_windows(*a, **k):
            raise RuntimeError("scikit-image is stubbed in validate_against_sitk.py (patch_correlation vote only)")

        sk_shape.view_as_windows = view_as_windows
        sys.modules.setdefault("skimage", sk_pkg)
        sys.modules.setdefault("skimage.util", sk_util)
        sys.modules.setdefault("skimage.util.shape", sk_shape)
    ns = types.SimpleNamespace()
    ns.deformable = importlib.import_module("platipy.imaging.registration.deformable")
    ns.utils = importlib.import_module("platipy.imaging.registration.utils")
    ns.fusion = importlib.import_module("platipy.imaging.label.fusion")
    ns.fast_symmetric_forces_demons_registration = ns.deformable.fast_symmetric_forces_demons_registration  # deformable.py:190
    ns.multiscale_demons = ns.deformable.multiscale_demons                                                  # deformable.py:31
    ns.apply_transform = ns.utils.apply_transform                                                           # utils.py:148
    ns.smooth_and_resample = ns.utils.smooth_and_resample                                                   # utils.py:195
    ns.compute_weight_map = ns.fusion.compute_weight_map                                                    # fusion.py:56
    ns.combine_labels = ns.fusion.combine_labels                                                            # fusion.py:239
    ns.combine_labels_staple = ns.fusion.combine_labels_staple                                              # fusion.py:205
    ns.process_probability_image = ns.fusion.process_probability_image                                      # fusion.py:295
    return ns


# ---------------------------------------------------------------------------------------------------------------------
# conversions between the stand-in Image of the package / oracle and SimpleITK images
# ---------------------------------------------------------------------------------------------------------------------
def to_sitk(img):
    import SimpleITK as sitk

    out = sitk.GetImageFromArray(img.array, isVector=bool(img.is_vector))
    out.SetSpacing(img.GetSpacing())
    out.SetOrigin(img.GetOrigin())
    out.SetDirection(img.GetDirection())
    return out


def arr(simg):
    import SimpleITK as sitk

    return sitk.GetArrayFromImage(simg)


class Report:
    def __init__(self):
        self.rows = []

    def add(self, name, diff, tol, note="", switch=None):
        row = {"check": name, "max_abs_diff": float(diff), "tolerance": float(tol), "pass": bool(diff <= tol), "note": note}
        if switch:
            row["switch"] = switch
        self.rows.append(row)
        print(("PASS " if row["pass"] else "FAIL ") + f"{name}: max |sitk - oracle| = {diff:.3e} (tol {tol:.1e}) {note}"
              + (f"  [{switch['name']}: matches SimpleITK at {switch['matches']}]" if switch else ""), flush=True)

    def ok(self):
        return all(r["pass"] for r in self.rows)


def _maxdiff(a, b):
    import numpy as np

    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    if a.shape != b.shape:
        return float("inf")
    return float(np.abs(a - b).max()) if a.size else 0.0


def with_switch(orc, name, fn):
    """Run ``fn()`` (-> difference to SimpleITK) under both settings of a switch; returns (diff at default, switch record)."""
    default = orc.get_semantic(name)
    d_default = fn()
    with orc.semantic(name, 1 - default):
        d_other = fn()
    matches = default if d_default <= d_other else 1 - default
    return d_default, {"name": name, "default": default, "diff_at_default": d_default, "diff_at_other": d_other, "matches": matches,
                       "default_is_right": bool(d_default <= d_other)}


def run_all(reference_root="/root/reference", gpu=False):
    import numpy as np
    import SimpleITK as sitk

    from oracle import itk_oracle as orc
    from oracle import platipy_ref as ref
    from platipy_b200 import sitk_compat as sk
    from platipy_b200.sitk_compat import Image
    from platipy_b200.synth import smooth_random_dvf, synth_labels, synth_pair

    P = import_reference(reference_root)
    rep = Report()
    print(f"SimpleITK {sitk.Version_VersionString()} (platipy locks 2.3.1), reference at {reference_root}")

    size, sp, org = (64, 64, 32), (1.0, 1.1, 1.5), (-10.0, 5.0, 20.0)
    fixed, moving = synth_pair(size, seed=0, spacing=sp, origin=org, moving_seed=100, peak_mm=3.0)
    sf, sm = to_sitk(fixed), to_sitk(moving)
    g = orc.geom_of(fixed)
    dvf = Image(smooth_random_dvf(size, seed=3, peak_mm=4.0), sp, org, is_vector=True)
    sdvf = to_sitk(dvf)
    other = Image(np.zeros((27, 41, 53), np.float32), (1.15, 1.4, 2.1), (-9.7, 5.2, 20.1))
    sother = to_sitk(other)
    labels = [Image(l, sp, org) for l in synth_labels(size, 3, seed=200)]

    # ---- GaussianOperator (A.5): coefficients through a delta image ---------------------------------------------------
    for var in (0.5, 1.0, 2.25, 16.0):
        delta = np.zeros((1, 1, 129), np.float32)
        delta[0, 0, 64] = 1.0
        k = arr(sitk.DiscreteGaussian(sitk.GetImageFromArray(delta), variance=[var, 0.0, 0.0], maximumKernelWidth=64, maximumError=0.01,
                                      useImageSpacing=False))[0, 0]
        ko = np.asarray(orc.gaussian_operator(var, 0.01, 64))
        r = (len(ko) - 1) // 2
        rep.add(f"GaussianOperator variance {var}", _maxdiff(k[64 - r:64 + r + 1], ko.astype(np.float32)), 1e-7, f"radius {r}")

    # ---- DiscreteGaussian (N1) + its pass order --------------------------------------------------------------------------
    want = arr(sitk.DiscreteGaussian(sf, 16.0, 128))
    d, sw = with_switch(orc, "discrete_gaussian_axis_order", lambda: _maxdiff(want, orc.discrete_gaussian_f32(fixed.array, g, [16.0] * 3, 128, 0.01, True)))
    rep.add("DiscreteGaussian variance 16 mm^2", d, REL_TOL * 1000, switch=sw)

    # ---- smooth_and_resample (utils.py:195-267) -----------------------------------------------------------------------------
    for kw in (dict(shrink_factor=2, smoothing_sigma=2.0), dict(isotropic_voxel_size_mm=3.0, smoothing_sigma=0), dict(shrink_factor=[4, 2, 1], smoothing_sigma=4.0)):
        a, b = P.smooth_and_resample(sf, **kw), ref.smooth_and_resample(fixed, **kw)
        rep.add(f"smooth_and_resample {kw}", max(_maxdiff(arr(a), b.array), _maxdiff(a.GetSpacing(), b.GetSpacing())), REL_TOL * 1000)

    # ---- apply_transform (utils.py:148-192): affine / field, linear / nearest neighbour, own grid / other grid --------------
    ang = 0.05
    m3 = [[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1]]
    saff = sitk.AffineTransform(3)
    saff.SetMatrix([v for r in m3 for v in r])
    saff.SetTranslation((0.7, -0.4, 0.3))
    saff.SetCenter((20.0, 25.0, 40.0))
    aff = sk.AffineTransform(m3, (0.7, -0.4, 0.3), (20.0, 25.0, 40.0))
    stfm = sitk.DisplacementFieldTransform(sitk.Image(sdvf))
    tfm = sk.DisplacementFieldTransform(dvf)
    d, sw = with_switch(orc, "resample_linear_scanline",
                        lambda: _maxdiff(arr(P.apply_transform(sm, sf, saff, -1000, sitk.sitkLinear)), ref.apply_transform(moving, fixed, aff, -1000, sk.sitkLinear).array))
    rep.add("apply_transform affine, linear", d, REL_TOL * 1000, switch=sw)
    rep.add("apply_transform affine, nearest neighbour (labels)", _maxdiff(arr(P.apply_transform(to_sitk(labels[0]), sf, saff, 0, sitk.sitkNearestNeighbor)),
                                                                           ref.apply_transform(labels[0], fixed, aff, 0, sk.sitkNearestNeighbor).array), 0.0, "bit-exact")
    rep.add("apply_transform field, linear, own grid", _maxdiff(arr(P.apply_transform(sm, sf, stfm, -1000, sitk.sitkLinear)),
                                                                ref.apply_transform(moving, fixed, tfm, -1000, sk.sitkLinear).array), REL_TOL * 1000)
    rep.add("apply_transform field, nearest neighbour (labels)", _maxdiff(arr(P.apply_transform(to_sitk(labels[1]), sf, stfm, 0, sitk.sitkNearestNeighbor)),
                                                                          ref.apply_transform(labels[1], fixed, tfm, 0, sk.sitkNearestNeighbor).array), 0.0, "bit-exact")
    d, sw = with_switch(orc, "dvf_transform_interpolation",
                        lambda: _maxdiff(arr(P.apply_transform(sm, sother, stfm, -1000, sitk.sitkLinear)), ref.apply_transform(moving, other, tfm, -1000, sk.sitkLinear).array))
    rep.add("apply_transform field, linear, other grid", d, REL_TOL * 1000, switch=sw)

    # ---- vector Resample (N3 / N7) and the recursive Gaussian (N8) -----------------------------------------------------------
    d, sw = with_switch(orc, "vector_resample_interpolation", lambda: _maxdiff(arr(sitk.Resample(sdvf, sother)), ref.resample(dvf, other).array))
    rep.add("Resample of a VectorFloat64 field onto another grid", d, DVF_TOL_MM, switch=sw)
    want = arr(sitk.SmoothingRecursiveGaussian(sdvf, (1.5, 1.5 / 1.1, 1.0)))
    d, sw = with_switch(orc, "recursive_gaussian_axis_order", lambda: _maxdiff(want, orc.recursive_gaussian_vec3(dvf.array, orc.geom_of(dvf), (1.5, 1.5 / 1.1, 1.0))))
    rep.add("SmoothingRecursiveGaussian of a VectorFloat64 field", d, DVF_TOL_MM, switch=sw)

    # ---- the Demons filter itself (deformable.py:244-257,143-149) -------------------------------------------------------------
    filt = sitk.FastSymmetricForcesDemonsRegistrationFilter()
    filt.SetNumberOfThreads(1)
    filt.SetSmoothUpdateField(True)
    filt.SetSmoothDisplacementField(True)
    filt.SetStandardDeviations([1.5 / s for s in sp])
    filt.SetNumberOfIterations(10)
    want = arr(filt.Execute(sf, sm))
    D, st = orc.demons_execute(fixed.array, g, moving.array, g, orc.demons_params([1.5 / s for s in sp], 10, smooth_update_field=True))
    rep.add("FastSymmetricForcesDemons filter, 10 iterations: field", _maxdiff(want, D), DVF_TOL_MM)
    rep.add("FastSymmetricForcesDemons filter: elapsed iterations", abs(filt.GetElapsedIterations() - st["elapsed_iterations"]), 0)
    rep.add("FastSymmetricForcesDemons filter: metric", abs(filt.GetMetric() - st["metric"]) / max(1.0, abs(st["metric"])), 1e-9, "relative")
    rep.add("FastSymmetricForcesDemons filter: RMS change", abs(filt.GetRMSChange() - st["rms_change"]), 1e-9)

    # ---- fast_symmetric_forces_demons_registration (deformable.py:190-306): cfg1 and a 3-level case with an early stop ---------
    cases = {"cfg1 [1] x [10], sigma 0": dict(resolution_staging=[1], iteration_staging=[10], smoothing_sigmas=[0]),
             "cfg1 default sigmas": dict(resolution_staging=[1], iteration_staging=[10]),
             "3 levels [4, 2, 1] x [40, 20, 10]": dict(resolution_staging=[4, 2, 1], iteration_staging=[40, 20, 10]),
             "isotropic [6, 3] mm x [20, 10]": dict(resolution_staging=[6, 3], iteration_staging=[20, 10], isotropic_resample=True, smoothing_sigmas=[0, 0])}
    for name, kw in cases.items():
        img, _, field = P.fast_symmetric_forces_demons_registration(sf, sm, ncores=1, **kw)
        img_o, _, field_o = ref.fast_symmetric_forces_demons_registration(fixed, moving, **kw)
        rep.add(f"fast_symmetric_forces_demons_registration {name}: DVF", _maxdiff(arr(field), field_o.array), DVF_TOL_MM, "mm")
        rep.add(f"fast_symmetric_forces_demons_registration {name}: image", _maxdiff(arr(img), img_o.array) / 1000.0, REL_TOL, "relative to 1000 HU")

    # ---- fusion (fusion.py:56-328) ---------------------------------------------------------------------------------------------
    atlas, satlas = {}, {}
    rng = np.random.default_rng(7)
    for a in range(4):
        lab = {f"S{k}": Image(np.roll(labels[k].array, (a % 3 - 1, a - 2, 1 - a), axis=(0, 1, 2)), sp, org) for k in range(2)}
        w = Image((0.5 + rng.random(size[::-1])).astype(np.float32), sp, org)
        atlas[f"{a}"] = {"DIR": dict(lab, **{"Weight Map": w})}
        satlas[f"{a}"] = {"DIR": dict({k: to_sitk(v) for k, v in lab.items()}, **{"Weight Map": to_sitk(w)})}
    got, exp = P.combine_labels(satlas, ["S0", "S1"]), ref.combine_labels(atlas, ["S0", "S1"])
    rep.add("combine_labels, weighted, 4 atlases", max(_maxdiff(arr(got[s]), exp[s].array) for s in exp), REL_TOL)
    sl = {a: {k: v for k, v in satlas[a]["DIR"].items() if k != "Weight Map"} for a in satlas}
    ol = {a: {k: v for k, v in atlas[a]["DIR"].items() if k != "Weight Map"} for a in atlas}
    got, exp = P.combine_labels_staple(sl), ref.combine_labels_staple(ol)
    rep.add("combine_labels_staple, 4 raters", max(_maxdiff(arr(got[s]), exp[s].array) for s in exp), 1e-9)
    lab16 = Image((np.random.default_rng(0).random(size[::-1]) * 3).astype(np.int16), sp, org)
    want = arr(sitk.BinaryThreshold(to_sitk(lab16), lowerThreshold=0.5))
    d, sw = with_switch(orc, "binary_threshold_in_pixel_type", lambda: _maxdiff(want, ref._binary_threshold(lab16.array, 0.5, 255)))
    rep.add("BinaryThreshold(lowerThreshold=0.5) of an Int16 label", d, 0.0, switch=sw)
    for vt, params in (("unweighted", None), ("local", {"sigma": 2.0, "epsilon": 1e-5, "normalise": False}),
                       ("block", {"factor": 1e12, "gain": 6, "blockSize": 3, "normalise": False})):
        a, b = P.compute_weight_map(sf, sm, vt, params), ref.compute_weight_map(fixed, moving, vt, params)
        rep.add(f"compute_weight_map {vt}", _maxdiff(arr(a), b.array) / max(1.0, float(np.abs(b.array).max())), REL_TOL, "relative")
    prob = exp["S0"]
    rep.add("process_probability_image", _maxdiff(arr(P.process_probability_image(to_sitk(prob), 0.5)), ref.process_probability_image(prob, 0.5).array), 0.0,
            "bit-exact")

    # ---- optionally the CUDA path against SimpleITK directly ---------------------------------------------------------------------
    if gpu:
        from platipy_b200 import registration as reg

        for name, kw in cases.items():
            _, _, field = P.fast_symmetric_forces_demons_registration(sf, sm, ncores=1, **kw)
            _, _, field_g = reg.fast_symmetric_forces_demons_registration(fixed, moving, **kw)
            rep.add(f"GPU vs SimpleITK, {name}: DVF", _maxdiff(arr(field), field_g.array), DVF_TOL_MM, "mm")
    return rep


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--reference", default=os.environ.get("PLATIPY_REFERENCE", "/root/reference"))
    ap.add_argument("--gpu", action="store_true", help="also compare the CUDA path with SimpleITK directly")
    ap.add_argument("--json", default=None, help="write the report here")
    args = ap.parse_args()
    if not have_simpleitk():
        msg = {"unavailable": "SimpleITK is not importable here; nothing was validated (parity stays unpinned)"}
        print(json.dumps(msg))
        return 3
    rep = run_all(args.reference, args.gpu)
    out = {"all_pass": rep.ok(), "checks": rep.rows}
    if args.json:
        json.dump(out, open(args.json, "w"), indent=1)
    print(json.dumps({"all_pass": rep.ok(), "n_checks": len(rep.rows), "failed": [r["check"] for r in rep.rows if not r["pass"]]}))
    return 0 if rep.ok() else 1


if __name__ == "__main__":
    sys.exit(main())
