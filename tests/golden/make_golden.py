"""Generates tests/golden/*.npz.  These fixtures are ORACLE-generated (oracle/itk_oracle.c): the reference cannot run
here (SimpleITK is not installable offline) and its tests hold no golden vectors for this path, so parity stays
"unpinned" (DESIGN.md section 5).  The fixtures pin the oracle against regressions and give the GPU tests a
reference that does not depend on the oracle being rebuilt identically.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import platipy_ref as ref  # noqa: E402
from platipy_b200 import sitk_compat as sk  # noqa: E402
from platipy_b200.sitk_compat import Image  # noqa: E402


def phantom(shape, seed):
    rng = np.random.default_rng(seed)
    z, y, x = np.meshgrid(*[np.arange(s, dtype=np.float64) for s in shape], indexing="ij")
    v = -1000.0 + 1000.0 / (1.0 + np.exp(-(1.0 - (((x - shape[2] / 2) / (0.4 * shape[2])) ** 2 + ((y - shape[1] / 2) / (0.4 * shape[1])) ** 2 +
                                                   ((z - shape[0] / 2) / (0.4 * shape[0])) ** 2)) * 20.0))
    for _ in range(6):
        c = rng.uniform(0.3, 0.7, 3) * np.array(shape)
        v += rng.uniform(-300, 600) * np.exp(-((z - c[0]) ** 2 + (y - c[1]) ** 2 + (x - c[2]) ** 2) / (2 * rng.uniform(2, 5) ** 2))
    return v


def main():
    shape = (20, 28, 32)
    spacing, origin = (0.9, 0.95, 2.5), (320.0, -52.0, 60.0)
    fixed = phantom(shape, 1)
    z, y, x = np.meshgrid(*[np.arange(s, dtype=np.float64) for s in shape], indexing="ij")
    from scipy.ndimage import map_coordinates

    moving = map_coordinates(fixed, [z + 0.8 * np.sin(x / 9.0), y + 1.2 * np.cos(z / 5.0), x + 1.5 * np.sin(y / 7.0)], order=1, mode="nearest")
    rng = np.random.default_rng(2)
    fixed = (fixed + rng.normal(0, 3, shape)).astype(np.float32)
    moving = (moving + rng.normal(0, 3, shape)).astype(np.float32)
    label = ((z - 10) ** 2 / 36 + (y - 14) ** 2 / 64 + (x - 15) ** 2 / 81 <= 1).astype(np.uint8)
    F, M, L = Image(fixed, spacing, origin), Image(moving, spacing, origin), Image(label, spacing, origin)
    kw = dict(resolution_staging=[2, 1], iteration_staging=[6, 4])
    reg, tfm, dvf = ref.fast_symmetric_forces_demons_registration(F, M, **kw)
    lab = ref.apply_transform(L, F, tfm, 0, sk.sitkNearestNeighbor)
    np.savez_compressed(os.path.join(HERE, "demons_small.npz"), fixed=fixed, moving=moving, label=label, spacing=np.array(spacing),
                        origin=np.array(origin), resolution_staging=np.array([2, 1]), iteration_staging=np.array([6, 4]),
                        dvf=dvf.array, registered=reg.array, warped_label=lab.array)
    print("wrote demons_small.npz", dvf.array.shape, float(np.abs(dvf.array).max()))


def main_rows():
    """Second fixture: the rows added around the Demons loop (B-spline resampling, process_probability_image,
    block weight map, label utilities, mean-squares sums of linear_registration)."""
    from oracle import itk_oracle as orc

    g = np.load(os.path.join(HERE, "demons_small.npz"))
    spacing, origin = tuple(g["spacing"]), tuple(g["origin"])
    F, M = Image(g["fixed"], spacing, origin), Image(g["moving"], spacing, origin)
    tfm = sk.DisplacementFieldTransform(Image(g["dvf"], spacing, origin, is_vector=True))
    bsp = ref.apply_transform(M, F, tfm, -1000, sk.sitkBSpline)
    rng = np.random.default_rng(5)
    import scipy.ndimage as ndi

    v = ndi.gaussian_filter(rng.standard_normal(F.array.shape), 2.0)
    prob = np.clip((v - v.min()) / (v.max() - v.min()) * 0.9, 0, 1).astype(np.float32)
    mask = ref.process_probability_image(Image(prob, spacing, origin), 0.45)
    block = ref.compute_weight_map(F, M, "block", {"factor": 1e12, "gain": 6, "blockSize": (2, 2, 1), "normalise": True})
    labs = {"A": Image((prob > 0.5).astype(np.uint8), spacing, origin), "B": Image((np.roll(prob, 3, axis=2) > 0.55).astype(np.uint8), spacing, origin)}
    fixed_overlap = ref.correct_volume_overlap(labs)
    roi = ref.label_to_roi(labs["A"], [1, 1, 2.5], return_as_list=True)
    a = np.array([[0.99, -0.04, 0.01], [0.04, 1.01, 0.0], [0.0, 0.02, 0.98]])
    b = np.array([1.5, -2.0, 0.7]) + np.array(origin) - a @ np.array(origin)
    acc = ref.linreg_meansq(F, M, a, b, np.eye(3), np.array(origin), stride=3)
    np.savez_compressed(os.path.join(HERE, "rows_small.npz"), bspline=bsp.array, prob=prob, mask=mask.array, block=block.array,
                        overlap_a=fixed_overlap["A"].array, overlap_b=fixed_overlap["B"].array, roi=np.array(roi), lin_matrix=a, lin_offset=b,
                        lin_acc=acc)
    print("wrote rows_small.npz", int(mask.array.sum()), roi, acc[:2])


PATCH_PARAMS = {"patch_window_mm": 12, "resampled_voxel_size_mm": 3, "correlation_function": lambda x: x + 1}


def main_rows3(write=True):
    """Third fixture: the rows of round 1's third session (distance map, contours, morphology, reg structure, a generated field,
    the patch-correlation vote, correlation / Mattes sums, surface metrics) on the label and images of demons_small.npz."""
    from oracle import comparison_ref as cref
    from oracle import generation_ref as gref
    from oracle import itk_oracle as orc

    g = np.load(os.path.join(HERE, "demons_small.npz"))
    spacing, origin = tuple(g["spacing"]), tuple(g["origin"])
    F, M, L = Image(g["fixed"], spacing, origin), Image(g["moving"], spacing, origin), Image(g["label"], spacing, origin)
    out = {}
    out["maurer"] = orc.signed_maurer_distance_map(L.array, spacing)
    out["contour"] = orc.label_contour(L.array, False)
    out["dilated"] = gref.binary_dilate(L, (2, 1, 1)).array
    out["eroded"] = gref.binary_erode(L, (2, 1, 1)).array
    out["reg_structure"] = gref.convert_mask_to_reg_structure(L, 3).array
    shifted, _, dvf = gref.generate_field_shift(L, (2.5, -1.9, 1.8), 2)
    out["shift_mask"], out["shift_dvf"] = shifted.array, dvf.array
    out["patch_weight"] = ref.compute_weight_map(F, M, "patch_correlation", PATCH_PARAMS).array
    a = np.array([[0.99, -0.04, 0.01], [0.04, 1.01, 0.0], [0.0, 0.02, 0.98]])
    b = np.array([1.5, -2.0, 0.7]) + np.array(origin) - a @ np.array(origin)
    out["lin_matrix"], out["lin_offset"] = a, b
    out["corr_sums"] = ref.linreg_correlation(F, M, a, b, np.eye(3), np.array(origin), stride=3)
    fb = (float(F.array.max()) - float(F.array.min())) / 46, float(F.array.min()) / ((float(F.array.max()) - float(F.array.min())) / 46) - 2
    mb = (float(M.array.max()) - float(M.array.min())) / 46, float(M.array.min()) / ((float(M.array.max()) - float(M.array.min())) / 46) - 2
    hist, count = ref.linreg_mattes(F, M, a, b, np.eye(3), np.array(origin), fb, mb, stride=3)
    out["mattes_bins"], out["mattes_hist"], out["mattes_count"] = np.array(fb + mb), hist, np.array(count)
    other = Image(out["dilated"], spacing, origin)
    sm = cref.compute_surface_metrics(L, other)
    out["surface_metric_names"] = np.array(sorted(sm))
    out["surface_metric_values"] = np.array([float(sm[k]) for k in sorted(sm)])
    out["apl"] = np.array([int(v) for v in cref.compute_apl(L, other, 1.0)])
    if write:
        np.savez_compressed(os.path.join(HERE, "rows3_small.npz"), **out)
        print("wrote rows3_small.npz", float(np.abs(out["maurer"]).max()), int(out["shift_mask"].sum()), out["surface_metric_values"])
    return out


if __name__ == "__main__":
    main()
    main_rows()
    main_rows3()
