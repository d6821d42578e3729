"""
Multi-atlas segmentation, one atlas per GPU.

The reference processes atlases strictly serially (platipy/imaging/projects/multiatlas/run.py:261-362:
for every atlas -> Demons registration -> propagate CT + S labels -> weight map) and then fuses
(run.py:364 ``combine_labels``; fusion.py:205 ``combine_labels_staple``).  Atlases are independent units, so
they are partitioned over the ranks of a ``torch.distributed`` process group (atlas ``a`` -> rank
``a mod world``) and the path's ONE exchange step is an all-reduce (sum) of the per-voxel vote volume:

    weighted vote   float32 [num = sum_a w_a L_a , den = sum_a w_a] per structure          (8 B/voxel/structure)
    STAPLE          int32 bit mask, bit a = decision of atlas a (sum == OR, bits are disjoint) (4 B/voxel/structure)

followed by the replicated finalisation (normalise -> DiscreteGaussian -> rescale -> threshold, or the
STAPLE EM) and ``process_probability_image``.  A single Demons registration is not sharded.  With
``settings["linear_registration_settings"]`` the atlases may live in their own space and are first aligned with
``linear_registration`` (run.py:261-300); without it they are expected on the target grid already (the state
after ``apply_transform`` with the rigid transform).  With ``auto_crop_target_image_settings`` the target is first cropped
to the region the atlases cover (run.py:200-259) and the results are pasted back into the full grid (run.py:387-404).

``shard_atlases`` / ``exchange_sum`` are plain host logic and are exercised on CPU with the gloo backend.
"""
from __future__ import annotations

import numpy as np

from . import sitk_compat as sk

MULTIATLAS_SETTINGS_DEFAULTS = {
    "deformable_registration_settings": {
        "isotropic_resample": True,
        "resolution_staging": [16, 8, 4],
        "iteration_staging": [5, 5, 5],
        "smoothing_sigmas": [0, 0, 0],
        "ncores": 8,
        "default_value": -1000,
        "verbose": False,
    },
    "label_fusion_settings": {
        "vote_type": "unweighted",
        "vote_params": None,
        "optimal_threshold": {},
        "fusion": "vote",  # "vote" = combine_labels (reference pipeline), "staple" = combine_labels_staple
    },
}


MUTLIATLAS_SETTINGS_DEFAULTS = MULTIATLAS_SETTINGS_DEFAULTS  # the reference's spelling of the name (multiatlas/run.py:47)


def shard_atlases(atlas_ids, rank, world_size):
    """Atlas ids handled by ``rank``: sorted ids, round-robin (atlas a -> rank a mod world)."""
    ids = sorted(atlas_ids)
    return [a for i, a in enumerate(ids) if i % world_size == rank]


def atlas_bit(atlas_ids, atlas_id):
    """Global bit index of an atlas in the STAPLE decision mask."""
    return sorted(atlas_ids).index(atlas_id)


def exchange_sum(tensors, group=None):
    """The path's one collective: in-place sum over ranks of every tensor in ``tensors`` (NCCL for CUDA
    tensors, gloo for CPU tensors).  No-op without an initialised process group."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return tensors
    for t in tensors:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return tensors


def _dist_info(group=None):
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def rigid_align_atlas(target, atlas_ct, atlas_labels, settings):
    """Step 2 of the reference pipeline for one atlas (run.py:261-300): linear_registration of the atlas CT to the
    target, then the CT (linear, -1000) and every structure (nearest neighbour, 0) through that transform onto the
    target grid, in one batched pass.  Returns ``(ct, {structure: label}, transform)`` as device images."""
    from . import linear
    from . import registration as reg
    from .engine import Engine

    eng = Engine.get()
    t, m = eng.to_device(target), eng.to_device(atlas_ct)
    kw = {k: v for k, v in settings["linear_registration_settings"].items()}
    _, tfm = linear.linear_registration(t, m, **kw)
    names = list(atlas_labels)
    imgs = [m] + [eng.to_device(atlas_labels[n]) for n in names]
    outs = reg.apply_transform_batch(imgs, t, tfm, [-1000] + [0] * len(names), [sk.sitkLinear] + [sk.sitkNearestNeighbor] * len(names))
    return outs[0], dict(zip(names, outs[1:])), tfm


def register_atlas(target, atlas_ct, atlas_labels, settings):
    """Steps 3 of the reference pipeline for one atlas (run.py:312-347), device resident: Demons registration
    of the atlas CT to the target, then the CT (linear, -1000) and every label (nearest neighbour, 0) through
    the one transform in a single batched pass.  Returns ``{"CT Image", <structures>..., "Transform"}``."""
    from . import registration as reg
    from .engine import Engine

    eng = Engine.get()
    t, m = eng.to_device(target), eng.to_device(atlas_ct)
    kw = dict(settings["deformable_registration_settings"])
    _, tfm, _ = reg.fast_symmetric_forces_demons_registration(t, m, **kw)
    names = list(atlas_labels)
    imgs = [m] + [eng.to_device(atlas_labels[n]) for n in names]
    outs = reg.apply_transform_batch(imgs, t, tfm, [-1000] + [0] * len(names), [sk.sitkLinear] + [sk.sitkNearestNeighbor] * len(names))
    out = {"CT Image": outs[0], "Transform": tfm}
    for n, o in zip(names, outs[1:]):
        out[n] = o
    return out


def load_atlas_set(settings):
    """Initialisation step of the reference (run.py:147-190): read every atlas image and structure named by
    ``settings["atlas_settings"]`` (``atlas_path``, ``atlas_id_list``, ``atlas_structure_list``, ``atlas_image_format``,
    ``atlas_label_format``) and, with ``crop_atlas_to_structures``, crop each atlas to the bounding box of its structures
    expanded by ``crop_atlas_expansion_mm``.  Returns ``{atlas_id: {"CT Image": image, structure: label, ...}}``."""
    from . import label_utils as lu

    a = settings["atlas_settings"]
    atlas_set = {}
    for atlas_id in a["atlas_id_list"]:
        image = sk.ReadImage(f"{a['atlas_path']}/{a['atlas_image_format'].format(atlas_id)}")
        structures = {s: sk.ReadImage(f"{a['atlas_path']}/{a['atlas_label_format'].format(atlas_id, s)}") for s in a["atlas_structure_list"]}
        if a.get("crop_atlas_to_structures", False):
            size, index = lu.label_to_roi(list(structures.values()), expansion_mm=a["crop_atlas_expansion_mm"])
            image = lu.crop_to_roi(image, size=size, index=index)
            structures = {s: lu.crop_to_roi(v, size=size, index=index) for s, v in structures.items()}
        entry = {"CT Image": image}
        entry.update(structures)
        atlas_set[atlas_id] = entry
    return atlas_set


def run_segmentation(img, atlas_set=None, settings=MULTIATLAS_SETTINGS_DEFAULTS, group=None):
    """``run_segmentation`` (run.py:106-441).  Called like the reference -- ``run_segmentation(img, settings)`` with an
    ``atlas_settings`` block -- the atlases are read from disk (``load_atlas_set``); an in-memory ``atlas_set`` may be
    passed instead.

    img        target image (host ``Image`` or ``DeviceImage``)
    atlas_set  ``{atlas_id: {"CT Image": image, "<structure>": label image, ...}}`` on the target grid; every rank
               passes the FULL dictionary (or at least its own shard) and processes ``shard_atlases(...)`` of it.
    returns    ``(results, results_prob)``: binary UInt8 masks and fused probability images per structure
               (host images on every rank).
    """
    import torch

    from . import fusion
    from .engine import Engine

    from . import label_utils as lu

    if isinstance(atlas_set, dict) and "atlas_settings" in atlas_set:  # the reference's positional form: (img, settings)
        atlas_set, settings = None, atlas_set
    if atlas_set is None:
        atlas_set = load_atlas_set(settings)
    eng = Engine.get()
    rank, world = _dist_info(group)
    all_ids = sorted(atlas_set)
    mine = shard_atlases(all_ids, rank, world)
    structures = sorted({k for a in all_ids for k in atlas_set[a] if k != "CT Image"})
    fs = settings["label_fusion_settings"]
    vote_type, vote_params = fs.get("vote_type", "unweighted"), fs.get("vote_params", None)
    full = eng.to_device(img)
    target, crop_box = full, None
    if settings.get("auto_crop_target_image_settings") and settings.get("linear_registration_settings"):
        # Step 1 (run.py:200-246): quick similarity registration of up to 8 atlases, bounding box of the mean registered
        # image > -1000 expanded by expansion_mm, crop the target.  Sharded: local sums, one all-reduce.
        from . import linear

        quick = {"reg_method": "similarity", "shrink_factors": [8], "smooth_sigmas": [0], "sampling_rate": 0.75, "default_value": -1000,
                 "number_of_iterations": 25, "final_interp": sk.sitkLinear, "metric": "mean_squares", "optimiser": "gradient_descent_line_search"}
        crop_ids = all_ids[: min(8, len(all_ids))]
        acc = eng.zeros(full.tensor.shape, np.float32)
        for a in crop_ids:
            if a in mine:
                reg_image, _ = linear.linear_registration(full, eng.to_device(atlas_set[a]["CT Image"]), **quick)
                with torch.cuda.stream(eng.stream):
                    acc += eng.cast(reg_image, np.float32).tensor
        with torch.cuda.stream(eng.stream):
            exchange_sum([acc], group)
            combined = full.like((acc / float(len(crop_ids)) > -1000).to(torch.uint8), np.uint8, False)
        crop_size, crop_index = lu.label_to_roi(combined, expansion_mm=settings["auto_crop_target_image_settings"]["expansion_mm"])
        target = lu.crop_to_roi(full, crop_size, crop_index)
        crop_box = (crop_size, crop_index)
    tgt_f32 = eng.cast(target, np.float32)
    z, y, x = target.tensor.shape

    # ---- per-atlas work (independent units) --------------------------------------------------------------
    local = {}
    for a in mine:
        labels = {k: v for k, v in atlas_set[a].items() if k != "CT Image"}
        ct = atlas_set[a]["CT Image"]
        if settings.get("linear_registration_settings"):
            # run.py:261-300: atlases arrive in their own space and are first aligned linearly
            ct, labels, _ = rigid_align_atlas(target, ct, labels, settings)
        d = register_atlas(target, ct, labels, settings)
        d["Weight Map"] = fusion.compute_weight_map(tgt_f32, eng.cast(d["CT Image"], np.float32), vote_type, vote_params)
        local[a] = {"DIR": d}

    # ---- the one exchange + replicated finalisation -------------------------------------------------------
    results_prob = {}
    if fs.get("fusion", "vote") == "staple":
        packed = {}
        for s in structures:
            acc = eng.zeros((z, y, x), np.int32)
            for a in mine:
                if s in local[a]["DIR"]:
                    lab = eng.binary_threshold(local[a]["DIR"][s], 0.5, 255.0)
                    eng.pack_decision(lab, atlas_bit(all_ids, a), acc, False)
            packed[s] = acc
        with torch.cuda.stream(eng.stream):
            exchange_sum(list(packed.values()), group)
        for s in structures:
            holders = [a for a in all_ids if s in atlas_set[a]]
            dec = [eng.unpack_decision(packed[s], atlas_bit(all_ids, a), target) for a in holders]
            prob, _ = eng.staple(dec, threshold=1e-4, rescale=True)
            results_prob[s] = prob
    else:
        nums, dens = {}, {}
        for s in structures:
            num, den, _ = fusion.accumulate_votes(eng, local, s, "DIR")
            if num is None:
                num, den = eng.zeros((z, y, x), np.float32), eng.zeros((z, y, x), np.float32)
            nums[s], dens[s] = num, den
        with torch.cuda.stream(eng.stream):
            exchange_sum(list(nums.values()) + list(dens.values()), group)
        for s in structures:
            results_prob[s] = eng.vote_finalize(nums[s], dens[s], target, 1.0, 1e-4)

    # ---- binary masks (run.py:370-404): process_probability_image on the device, pasted back into the uncropped grid ----
    masks = {}
    for s in structures:
        thr = fs.get("optimal_threshold", {}).get(s, 0.5)
        masks[s] = eng.process_probability(results_prob[s], thr)
        if crop_box is not None:
            with torch.cuda.stream(eng.stream):
                tmpl_b = full.like(torch.zeros(full.tensor.shape, dtype=torch.uint8, device=full.tensor.device), np.uint8, False)
                tmpl_p = full.like(torch.zeros(full.tensor.shape, dtype=results_prob[s].tensor.dtype, device=full.tensor.device),
                                   results_prob[s].np_dtype, False)
            masks[s] = eng.region_copy(masks[s], (0, 0, 0), tmpl_b, crop_box[1], crop_box[0])
            results_prob[s] = eng.region_copy(results_prob[s], (0, 0, 0), tmpl_p, crop_box[1], crop_box[0])

    # ---- post-processing (run.py:409-437) ----------------------------------------------------------------------------
    pp = settings.get("postprocessing_settings") or {}
    if pp.get("run_postprocessing"):
        radius = [int(pp["binaryfillhole_mm"] / sp) for sp in full.GetSpacing()]
        for s in pp.get("structures_for_binaryfillhole", []):
            if s not in masks:
                continue
            # sitk.RelabelComponent(sitk.ConnectedComponent(x)) == 1, then BinaryMorphologicalClosing
            masks[s] = eng.binary_closing(eng.largest_component(masks[s]), radius, lu.ball_offsets(radius))
        oc = [s for s in pp.get("structures_for_overlap_correction", [])]
        if len(oc) >= 2:
            fixed = lu.correct_volume_overlap({s: masks[s] for s in oc})
            for s in oc:
                masks[s] = fixed[s]

    results, probs_host = {}, {}
    for s in structures:
        results[s] = eng.to_host(masks[s])
        probs_host[s] = eng.to_host(results_prob[s])
    return results, probs_host
