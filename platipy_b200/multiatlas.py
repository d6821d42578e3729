"""
Multi-atlas segmentation, atlases sharded over GPUs and fusion sharded over structures.

The reference processes atlases strictly serially (platipy/imaging/projects/multiatlas/run.py:261-362:
for every atlas -> linear registration -> Demons registration -> propagate CT + S labels -> weight map) and then
fuses structure by structure (run.py:364 ``combine_labels``; fusion.py:205 ``combine_labels_staple``), thresholds and
post-processes every structure (run.py:370-437).  Both loops are over independent units:

  * atlases are partitioned over the ranks of a ``torch.distributed`` process group (atlas ``a`` -> rank ``a mod world``);
  * structures are partitioned the same way for the fusion tail (structure ``s`` -> rank ``s mod world``): the path's ONE
    data exchange is a reduce-scatter (sum) over the structure axis of the per-voxel vote volume, after which every rank
    finalises only the structures it owns (STAPLE EM / normalise -> DiscreteGaussian -> rescale -> threshold, then
    ``process_probability_image``, paste, per-structure post-processing) and the UInt8 masks are all-gathered.

Exchange payload per voxel and structure:

    STAPLE            decision mask, bit a = decision of atlas a: 1 byte for <= 8 atlases (2 for <= 16).  Ranks own disjoint
                      bits, so the byte-wise SUM of the reduce-scatter is the bitwise OR, and the reduced mask *is* the
                      decision pattern the pattern-histogram EM works on (no unpacking).
    unweighted vote   UInt8 count of the atlases voting 1 (the float32 sums of the reference are exact small integers);
                      the denominator -- the number of atlases holding the structure -- is known on the host.
    weighted vote     float32 sum_a w_a L_a per structure (reduce-scatter) + float32 sum_a w_a per distinct set of holder
                      atlases (all-reduce; one volume when every atlas has every structure).

A single Demons registration is not sharded.  With ``settings["linear_registration_settings"]`` the atlases may live in
their own space and are first aligned with ``linear_registration`` (run.py:261-300); without it they must already be on the
target grid.  With ``auto_crop_target_image_settings`` the target is first cropped to the region the atlases cover
(run.py:200-259) and the results are pasted back into the full grid (run.py:387-404).

``shard_atlases`` / ``shard_structures`` / ``exchange_*`` are plain host logic and are exercised on CPU with the gloo backend.
"""
from __future__ import annotations

import copy
import os

import numpy as np

from . import sitk_compat as sk

ATLAS_PATH = os.environ.get("ATLAS_PATH", "/atlas")  # run.py:42-44

# The reference's defaults, key for key (run.py:47-103).  ``label_fusion_settings["fusion"]`` ("vote" = combine_labels, the
# reference pipeline; "staple" = combine_labels_staple) is this package's one addition and defaults to "vote" when absent.
MUTLIATLAS_SETTINGS_DEFAULTS = {
    "atlas_settings": {
        "atlas_id_list": ["03"],
        "atlas_structure_list": ["WHOLEHEART"],
        "atlas_path": ATLAS_PATH,
        "atlas_image_format": "Case_{0}/Images/Case_{0}_CROP.nii.gz",
        "atlas_label_format": "Case_{0}/Structures/Case_{0}_{1}_CROP.nii.gz",
        "crop_atlas_to_structures": False,
        "crop_atlas_expansion_mm": (20, 20, 40),
    },
    "auto_crop_target_image_settings": {"expansion_mm": [20, 20, 40]},
    "linear_registration_settings": {
        "reg_method": "affine",
        "shrink_factors": [16, 8, 4],
        "smooth_sigmas": [0, 0, 0],
        "sampling_rate": 0.75,
        "default_value": None,
        "number_of_iterations": 50,
        "metric": "mean_squares",
        "optimiser": "gradient_descent_line_search",
        "verbose": False,
    },
    "deformable_registration_settings": {
        "isotropic_resample": True,
        "resolution_staging": [6, 3, 1.5],  # voxel sizes (mm) since isotropic_resample is set
        "iteration_staging": [150, 125, 100],
        "smoothing_sigmas": [0, 0, 0],
        "ncores": 8,
        "default_value": None,
        "verbose": False,
    },
    "label_fusion_settings": {"vote_type": "unweighted", "vote_params": None, "optimal_threshold": {}},
    "postprocessing_settings": {
        "run_postprocessing": True,
        "binaryfillhole_mm": 3,
        "structures_for_binaryfillhole": [],
        "structures_for_overlap_correction": [],
    },
}
MULTIATLAS_SETTINGS_DEFAULTS = MUTLIATLAS_SETTINGS_DEFAULTS  # the reference spells it MUTLIATLAS (run.py:47); both names, one dict

# Short settings for atlases that already live on the target grid (tests, smoke runs): Demons only, no linear step, no crop.
ON_GRID_QUICK_SETTINGS = {
    "deformable_registration_settings": {
        "isotropic_resample": True,
        "resolution_staging": [16, 8, 4],
        "iteration_staging": [5, 5, 5],
        "smoothing_sigmas": [0, 0, 0],
        "ncores": 8,
        "default_value": -1000,
        "verbose": False,
    },
    "label_fusion_settings": {"vote_type": "unweighted", "vote_params": None, "optimal_threshold": {}, "fusion": "vote"},
}


# ---------------------------------------------------------------------------------------------------------------------
# partitioning and the exchange (host logic; CPU-tested with gloo)
# ---------------------------------------------------------------------------------------------------------------------
def shard_atlases(atlas_ids, rank, world_size):
    """Atlas ids handled by ``rank``: sorted ids, round-robin (atlas a -> rank a mod world)."""
    ids = sorted(atlas_ids)
    return [a for i, a in enumerate(ids) if i % world_size == rank]


def atlas_bit(atlas_ids, atlas_id):
    """Global bit index of an atlas in the STAPLE decision mask."""
    return sorted(atlas_ids).index(atlas_id)


def shard_structures(structures, world_size):
    """Structure ownership for the fusion tail: ``(per, slots)`` with ``slots[r]`` the ``per`` structure names owned by rank
    ``r`` (round-robin over the given order, padded with ``None``) -- the layout of the reduce-scatter stack, rank-major."""
    structures = list(structures)
    per = max(1, -(-len(structures) // world_size))
    slots = [[None] * per for _ in range(world_size)]
    for i, s in enumerate(structures):
        slots[i % world_size][i // world_size] = s
    return per, slots


def mask_dtype(n_atlases):
    """torch container of the STAPLE decision mask: 1 byte up to 8 atlases, 2 up to 16, 4 beyond."""
    import torch

    return torch.uint8 if n_atlases <= 8 else (torch.int16 if n_atlases <= 16 else torch.int32)


def _dist_info(group=None):
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def exchange_sum(tensors, group=None):
    """In-place sum over ranks of every tensor in ``tensors`` (NCCL for CUDA tensors, gloo for CPU tensors).  No-op without
    an initialised process group.  Used for the auto-crop accumulator and the weighted vote's denominators."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return tensors
    for t in tensors:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return tensors


def _lane_views(stack, out):
    """Integer payloads travel in the widest integer type their size allows.  Decision masks: ranks set disjoint bits, so a SUM never
    carries and equals the OR whatever the lane width (NCCL has no 16-bit integer type anyway).  Vote counts: every byte stays
    <= 255 over all ranks (at most 255 atlases, labels in {0, 1}), so byte lanes never carry into their neighbours either -- summing
    eight of them as one int64 gives the same bytes with an eighth of the elements.  Returns flat views of ``stack`` and ``out``
    (``out`` is one rank's equal share of ``stack``, or ``stack`` itself) in the common lane type."""
    import torch

    if stack.dtype == torch.float32:
        return stack, out
    a, b = stack.reshape(-1).view(torch.uint8), out.reshape(-1).view(torch.uint8)
    if b.numel() % 8 == 0 and a.data_ptr() % 8 == 0 and b.data_ptr() % 8 == 0:
        return a.view(torch.int64), b.view(torch.int64)
    return a, b


def exchange_reduce_scatter(stack, group=None):
    """The path's one data exchange.  ``stack``: ``[world * per, z, y, x]``, rank-major (``shard_structures``); returns the
    ``[per, z, y, x]`` block of this rank summed over all ranks.  NCCL: one ``reduce_scatter_tensor``; gloo (CPU tests) has no
    reduce-scatter, there it is an all-reduce followed by the slice."""
    import torch
    import torch.distributed as dist

    rank, world = _dist_info(group)
    if world == 1:
        return stack
    per = stack.shape[0] // world
    if dist.get_backend(group) == "nccl":
        out = torch.empty((per,) + tuple(stack.shape[1:]), dtype=stack.dtype, device=stack.device)
        vin, vout = _lane_views(stack, out)
        dist.reduce_scatter_tensor(vout, vin, op=dist.ReduceOp.SUM, group=group)
        return out
    vin, _ = _lane_views(stack, stack)
    dist.all_reduce(vin, op=dist.ReduceOp.SUM, group=group)  # gloo (CPU tests): no reduce-scatter
    return stack[rank * per:(rank + 1) * per]


def exchange_all_gather(block, group=None):
    """``[per, ...]`` per rank -> ``[world * per, ...]`` on every rank, rank-major."""
    import torch
    import torch.distributed as dist

    rank, world = _dist_info(group)
    if world == 1:
        return block
    out = torch.empty((world * block.shape[0],) + tuple(block.shape[1:]), dtype=block.dtype, device=block.device)
    if dist.get_backend(group) == "nccl":
        dist.all_gather_into_tensor(out, block.contiguous(), group=group)
    else:
        parts = [torch.empty_like(block) for _ in range(world)]
        dist.all_gather(parts, block.contiguous(), group=group)
        out = torch.cat(parts, dim=0)
    return out


def _gather_holders(local_holders, group=None):
    """{atlas id: [structure names]} of every rank's local atlases -> the global dictionary (host metadata only)."""
    import torch.distributed as dist

    rank, world = _dist_info(group)
    if world == 1:
        return dict(local_holders)
    parts = [None] * world
    dist.all_gather_object(parts, local_holders, group=group)
    out = {}
    for p in parts:
        out.update(p)
    return out


# ---------------------------------------------------------------------------------------------------------------------
# per-atlas steps (run.py:261-347)
# ---------------------------------------------------------------------------------------------------------------------
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
    if t.GetSize() != m.GetSize():
        # ITK: "Inputs do not occupy the same physical space!" -- the reference reaches Demons only after the linear step
        raise RuntimeError("run_segmentation: the atlas image is not on the target grid; give settings['linear_registration_settings'] "
                           "(run.py:261-300 aligns and resamples every atlas onto the target before Demons)")
    kw = dict(settings["deformable_registration_settings"])
    _, tfm, _ = reg.fast_symmetric_forces_demons_registration(t, m, **kw)
    names = list(atlas_labels)
    imgs = [m] + [eng.to_device(atlas_labels[n]) for n in names]
    outs = reg.apply_transform_batch(imgs, t, tfm, [-1000] + [0] * len(names), [sk.sitkLinear] + [sk.sitkNearestNeighbor] * len(names))
    out = {"CT Image": outs[0], "Transform": tfm}
    for n, o in zip(names, outs[1:]):
        out[n] = o
    return out


def load_atlas_set(settings, only=None):
    """Initialisation step of the reference (run.py:147-190): read every atlas image and structure named by
    ``settings["atlas_settings"]`` (``atlas_path``, ``atlas_id_list``, ``atlas_structure_list``, ``atlas_image_format``,
    ``atlas_label_format``) and, with ``crop_atlas_to_structures``, crop each atlas to the bounding box of its structures
    expanded by ``crop_atlas_expansion_mm``.  ``only``: read just these ids (a rank reads its own shard).
    Returns ``{atlas_id: {"CT Image": image, structure: label, ...}}``."""
    from . import label_utils as lu

    a = settings["atlas_settings"]
    atlas_set = {}
    for atlas_id in a["atlas_id_list"]:
        if only is not None and atlas_id not in only:
            continue
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


# ---------------------------------------------------------------------------------------------------------------------
# the pipeline
# ---------------------------------------------------------------------------------------------------------------------
def run_segmentation(img, atlas_set=None, settings=MUTLIATLAS_SETTINGS_DEFAULTS, group=None, atlas_ids=None, gather_probabilities=True,
                     timings=None):
    """``run_segmentation`` (run.py:106-441).  Called like the reference -- ``run_segmentation(img, settings)`` with an
    ``atlas_settings`` block -- the atlases are read from disk (``load_atlas_set``; under a process group every rank reads
    its own shard); an in-memory ``atlas_set`` may be passed instead.

    img        target image (host ``Image`` or ``DeviceImage``; a device target gives device results)
    atlas_set  ``{atlas_id: {"CT Image": image, "<structure>": label image, ...}}``.  Under a process group either every rank
               passes the full dictionary, or every rank passes (at least) its own shard together with ``atlas_ids``, the
               full id list, identical on all ranks: shards and STAPLE bits are derived from that list only.
    gather_probabilities  under a process group, whether the fused probability maps of all structures are gathered on every
               rank (the reference's return value) or every rank returns the maps of the structures it owns only.
    timings    optional dict filled with host-clock seconds per stage (each stage is synchronised when given).
    returns    ``(results, results_prob)``: binary UInt8 masks (all structures, every rank) and fused probability images.
    """
    import time

    import torch

    from . import fusion
    from . import label_utils as lu
    from .engine import DeviceImage, Engine

    if isinstance(atlas_set, dict) and "atlas_settings" in atlas_set:  # the reference's positional form: (img, settings)
        atlas_set, settings = None, atlas_set
    eng = Engine.get()
    rank, world = _dist_info(group)
    device_out = isinstance(img, DeviceImage)

    def tick(name, t0):
        if timings is not None:
            eng.synchronize()
            timings[name] = timings.get(name, 0.0) + time.perf_counter() - t0
        return time.perf_counter()

    t0 = time.perf_counter()
    if atlas_set is None:
        all_ids = sorted(settings["atlas_settings"]["atlas_id_list"])
        mine = shard_atlases(all_ids, rank, world)
        atlas_set = load_atlas_set(settings, only=mine)
    else:
        all_ids = sorted(atlas_ids) if atlas_ids is not None else sorted(atlas_set)
        mine = shard_atlases(all_ids, rank, world)
        missing = [a for a in mine if a not in atlas_set]
        if missing:
            raise KeyError(f"rank {rank}: atlases {missing} of its shard are not in atlas_set (pass the full id list as atlas_ids on every rank)")
    n_atlases = len(all_ids)
    # which atlas holds which structure, globally (host metadata; only exchanged when ranks hold partial dictionaries)
    local_holders = {a: sorted(k for k in atlas_set[a] if k != "CT Image") for a in atlas_set if a in all_ids}
    holders_of = local_holders if all(a in local_holders for a in all_ids) else _gather_holders({a: local_holders[a] for a in mine}, group)
    structures = sorted({k for a in all_ids for k in holders_of[a]})
    fs = settings["label_fusion_settings"]
    vote_type, vote_params = fs.get("vote_type", "unweighted"), fs.get("vote_params", None)
    mode = fs.get("fusion", "vote")
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
    t0 = tick("setup_s", t0)

    # ---- per-atlas work (independent units) --------------------------------------------------------------
    unweighted = vote_type.lower() == "unweighted"
    local = {}
    for a in mine:
        labels = {k: v for k, v in atlas_set[a].items() if k != "CT Image"}
        ct = atlas_set[a]["CT Image"]
        if settings.get("linear_registration_settings"):
            # run.py:261-300: atlases arrive in their own space and are first aligned linearly
            ct, labels, _ = rigid_align_atlas(target, ct, labels, settings)
            t0 = tick("linear_s", t0)
        d = register_atlas(target, ct, labels, settings)
        t0 = tick("deformable_s", t0)
        if mode != "staple" and not unweighted:
            d["Weight Map"] = fusion.compute_weight_map(tgt_f32, eng.cast(d["CT Image"], np.float32), vote_type, vote_params)
        local[a] = {"DIR": d}
        t0 = tick("weight_map_s", t0)

    # ---- the one exchange: reduce-scatter over the structure axis ------------------------------------------
    per, slots = shard_structures(structures, world)
    my_slots = slots[rank]
    holder_ids = {s: [a for a in all_ids if s in holders_of[a]] for s in structures}
    dens = {}
    if mode == "staple":
        if n_atlases > 16:
            raise NotImplementedError("the packed STAPLE exchange handles up to 16 atlases; use combine_labels_staple for more")
        with torch.cuda.stream(eng.stream):
            stack = torch.zeros((world * per, z, y, x), dtype=mask_dtype(n_atlases), device=eng.device)
        for r in range(world):
            for j, s in enumerate(slots[r]):
                for a in mine:
                    if s is not None and s in local[a]["DIR"]:
                        lab = local[a]["DIR"][s]
                        if lab.np_dtype != np.uint8:  # BinaryThreshold(lowerThreshold=0.5) (fusion.py:217-220); for UInt8 it is "!= 0"
                            lab = eng.binary_threshold(lab, 0.5, 255.0)
                        eng.pack_label(lab, atlas_bit(all_ids, a), stack[r * per + j], False)
        payload = "mask"
    elif unweighted and n_atlases <= 255:
        with torch.cuda.stream(eng.stream):
            stack = torch.zeros((world * per, z, y, x), dtype=torch.uint8, device=eng.device)
            flag = torch.zeros(1, dtype=torch.int32, device=eng.device)
        for r in range(world):
            for j, s in enumerate(slots[r]):
                for a in mine:
                    if s is not None and s in local[a]["DIR"]:
                        eng.count_accumulate(eng.cast(local[a]["DIR"][s], np.uint8), stack[r * per + j], False, flag)
        with torch.cuda.stream(eng.stream):
            exchange_sum([flag], group)
            binary = int(flag.item()) == 0
        payload = "count"
        if not binary:  # label values above 1: float32 sums of the label values, as the reference forms them
            payload = "float"
    else:
        payload = "float"
    if payload == "float":
        with torch.cuda.stream(eng.stream):
            stack = torch.zeros((world * per, z, y, x), dtype=torch.float32, device=eng.device)
        sets = {}
        for s in structures:
            sets.setdefault(tuple(holder_ids[s]), []).append(s)
        for hs in sets:  # one denominator volume per distinct set of holder atlases
            den = eng.zeros((z, y, x), np.float32)
            first = True
            for a in mine:
                if a in hs:
                    w = local[a]["DIR"].get("Weight Map")
                    if w is None:
                        w = local[a]["DIR"]["Weight Map"] = fusion.compute_weight_map(tgt_f32, eng.cast(local[a]["DIR"]["CT Image"], np.float32), vote_type, vote_params)
                    with torch.cuda.stream(eng.stream):
                        den = w.tensor.clone() if first else den + w.tensor  # plumbing: float32 adds in atlas order
                    first = False
            for s in sets[hs]:
                dens[s] = den
        for r in range(world):
            for j, s in enumerate(slots[r]):
                first = True
                for a in mine:
                    if s is not None and s in local[a]["DIR"]:
                        eng.vote_accumulate(eng.cast(local[a]["DIR"][s], np.uint8), local[a]["DIR"]["Weight Map"], stack[r * per + j], None, first)
                        first = False
    if timings is not None:
        timings.update(exchange_bytes=int(stack.numel() * stack.element_size()), grid=[int(x), int(y), int(z)], payload=payload)
    t0 = tick("pack_s", t0)
    with torch.cuda.stream(eng.stream):
        block = exchange_reduce_scatter(stack, group)
        if payload == "float":
            seen = set()
            for s in structures:
                if id(dens[s]) not in seen:
                    seen.add(id(dens[s]))
                    exchange_sum([dens[s]], group)
    del stack
    t0 = tick("exchange_s", t0)

    # ---- finalisation of the structures this rank owns (fusion.py:205-292, run.py:370-437) ----------------------------
    pp = settings.get("postprocessing_settings") or {}
    probs_mine = {}
    with torch.cuda.stream(eng.stream):
        mask_block = torch.zeros((per,) + tuple(full.tensor.shape), dtype=torch.uint8, device=eng.device)
    for j, s in enumerate(my_slots):
        if s is None:
            continue
        if payload == "mask":
            hm = sum(1 << atlas_bit(all_ids, a) for a in holder_ids[s])
            prob = eng.staple_packed(block[j], hm, target, threshold=1e-4, rescale=True)
        elif payload == "count":
            prob = eng.vote_finalize_counts(block[j], len(holder_ids[s]), target, 1.0, 1e-4)
        else:
            prob = eng.vote_finalize(block[j], dens[s], target, 1.0, 1e-4)
        thr = fs.get("optimal_threshold", {}).get(s, 0.5)
        mask = eng.process_probability(prob, thr)  # run.py:383
        if crop_box is not None:
            # run.py:387-404: paste into the uncropped grid
            with torch.cuda.stream(eng.stream):
                tmpl_b = full.like(torch.zeros(full.tensor.shape, dtype=torch.uint8, device=eng.device), np.uint8, False)
                tmpl_p = full.like(torch.zeros(full.tensor.shape, dtype=prob.tensor.dtype, device=eng.device), prob.np_dtype, False)
            mask = eng.region_copy(mask, (0, 0, 0), tmpl_b, crop_box[1], crop_box[0])
            prob = eng.region_copy(prob, (0, 0, 0), tmpl_p, crop_box[1], crop_box[0])
        if pp.get("run_postprocessing") and s in pp.get("structures_for_binaryfillhole", []):
            # run.py:421-431: sitk.RelabelComponent(sitk.ConnectedComponent(x)) == 1, then BinaryMorphologicalClosing
            radius = [int(pp["binaryfillhole_mm"] / sp) for sp in full.GetSpacing()]
            mask = eng.binary_closing(eng.largest_component(mask), radius, lu.ball_offsets(radius))
        with torch.cuda.stream(eng.stream):
            mask_block[j].copy_(mask.tensor)
        probs_mine[s] = prob
    del block
    t0 = tick("finalise_s", t0)

    # ---- gather the UInt8 masks; overlap correction needs all of them (run.py:433-441) -----------------------------
    with torch.cuda.stream(eng.stream):
        all_masks = exchange_all_gather(mask_block, group)
    masks = {}
    for r in range(world):
        for j, s in enumerate(slots[r]):
            if s is not None:
                masks[s] = full.like(all_masks[r * per + j], np.uint8, False)
    if pp.get("run_postprocessing"):
        oc = [s for s in pp.get("structures_for_overlap_correction", [])]
        if len(oc) >= 2:
            fixed = lu.correct_volume_overlap({s: masks[s] for s in oc})
            for s in oc:
                masks[s] = fixed[s]
    results_prob = dict(probs_mine)
    if world > 1 and gather_probabilities:
        pdt = torch.float64 if mode == "staple" else torch.float32
        with torch.cuda.stream(eng.stream):
            pblock = torch.zeros((per,) + tuple(full.tensor.shape), dtype=pdt, device=eng.device)
            for j, s in enumerate(my_slots):
                if s is not None:
                    pblock[j].copy_(probs_mine[s].tensor)
            allp = exchange_all_gather(pblock, group)
        results_prob = {}
        for r in range(world):
            for j, s in enumerate(slots[r]):
                if s is not None:
                    results_prob[s] = full.like(allp[r * per + j], np.float64 if mode == "staple" else np.float32, False)
    t0 = tick("gather_s", t0)

    if device_out:
        eng.release_to_caller()
        return {s: masks[s] for s in structures}, {s: results_prob[s] for s in structures if s in results_prob}
    results, probs_host = {}, {}
    for s in structures:
        results[s] = eng.to_host(masks[s])
        if s in results_prob:
            probs_host[s] = eng.to_host(results_prob[s])
    tick("to_host_s", t0)
    return results, probs_host


def settings_with(base=MUTLIATLAS_SETTINGS_DEFAULTS, **blocks):
    """Deep copy of a settings dictionary with whole blocks replaced (``None`` removes a block)."""
    out = copy.deepcopy(base)
    for k, v in blocks.items():
        if v is None:
            out.pop(k, None)
        else:
            out[k] = v
    return out
