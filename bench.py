#!/usr/bin/env python
"""
bench.py -- the reference's headline metric on B200 (BASELINE.json): Demons throughput at 512x512x256 and the multi-atlas
fusion wall-clock at 1 / 2 / 4 / 8 GPUs.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--size X Y Z] [--no-fusion] [--no-cpu-baseline]

Headline line (metric / value / e2e / roofline / cpu_baseline), BASELINE.json configs[1]:
  A "step" is one `fast_symmetric_forces_demons_registration` of a synthetic 512x512x256 float32 pair, 3-level pyramid
  [4, 2, 1], 100/50/25 iterations (early stop as in the reference).  N > 1: one pair per rank (rank r registers atlas 100+r to
  the common target: independent units, weak scaling, no data-path collective inside the step).
  metric   demons_voxel_iterations_per_s, unit Mvoxel*it/s: (sum over levels of voxels x elapsed iterations, summed over ranks) /
           (max over ranks of the device time of the K timed steps / K).
  value    inputs already resident in HBM (device-in, device-out call).
  e2e      the same metric through the public host API: pinned host buffers in, host images out; the H2D and D2H copies are
           inside the timed region, all K steps.
  roofline the full-resolution Demons iteration (the kernels between two iterations of level 2), algorithmic
           176 B/voxel/iteration with f64 fields, against MEASURED_PEAKS.json hbm_gbs.
  cpu_baseline / --impl reference: the CPU oracle (the restatement of the ITK filters; SimpleITK itself is not installable
           offline) on all host cores.  --impl reference times the WHOLE registration (pyramid, three levels, level glue, final
           warp) once and reports its throughput; the remaining steps are bounded samples of full-resolution iterations.

"fusion" (same JSON line; BASELINE.json configs[3] and configs[4], strong scaling -- total work fixed as N grows):
  `run_segmentation` (linear_registration -> Demons -> label propagation -> fusion -> process_probability_image, auto-crop and
  paste back) over a synthetic cohort: cfg4 = 4 atlases x 5 structures at 256x256x160, unweighted vote; cfg5 = 8 atlases x
  20 structures at 512x512x256, STAPLE.  Atlases are sharded over the N ranks, the fusion tail over structures; wall-clock is
  the max over ranks between two barriers, inputs resident in HBM on the rank that owns them, masks on every rank at the end.
  `stages_ms` (one extra instrumented run, every stage synchronised) separates per-atlas work, the exchange and the tail;
  `mask_checksums` are integer checksums of the final masks -- identical at every N when the sharded run is bit-exact.

"resample_cfg3" (same line; BASELINE.json configs[2]): apply_transform of one CT (linear) + 20 masks (nearest neighbour)
  through a dense f64 DVF at 512x512x256, per call and batched, with its own HBM roofline figures.

"platipy_default_staging" (same line, N = 1): the headline pair registered with the reference's default arguments ([8, 4, 1] x 10
  iterations); "fast_mode": precision="fast" beside parity mode with the error percentiles of the fast field (not a parity path).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

RES_STAGING = [4, 2, 1]
ITER_STAGING = [100, 50, 25]
BYTES_PER_VOXEL_ITER_F64 = 176.0  # force 56 + smooth-U 48 + add+smooth-D 72
METRIC = "demons_voxel_iterations_per_s"
UNIT = "Mvoxel*it/s"


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:  # noqa: BLE001
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:  # noqa: BLE001
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def bind_to_gpu_numa_node(local_rank):
    """Pin this rank's threads (and with them the first-touch placement of its pinned buffers) to the NUMA node its GPU hangs
    off: eight ranks copying 1.9 GB each per step otherwise contend for one node's memory controllers.  Best effort."""
    try:
        import torch

        pr = torch.cuda.get_device_properties(local_rank)
        if all(hasattr(pr, k) for k in ("pci_domain_id", "pci_bus_id", "pci_device_id")):
            bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        else:
            uuid = str(getattr(pr, "uuid", ""))
            out = subprocess.run(["nvidia-smi", "--query-gpu=uuid,pci.bus_id", "--format=csv,noheader"], capture_output=True, text=True, timeout=10).stdout
            rows = [[f.strip() for f in r.split(",")] for r in out.strip().splitlines()]
            match = [r[1] for r in rows if uuid and uuid in r[0]] or [rows[local_rank][1]]
            bdf = match[0].lower()
            if len(bdf.split(":")[0]) == 8:  # nvidia-smi prints an 8-digit PCI domain, sysfs uses 4
                bdf = bdf[4:]
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip())
        if node < 0:
            return None
        cpus = []
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus += list(range(int(a), int(b or a) + 1))
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
            return {"numa_node": node, "cpus": len(allowed)}
    except Exception:  # noqa: BLE001
        return None
    return None


CPU_SAMPLE_ITERS = 10  # full-resolution iterations per bounded CPU sample: 10-25 s on 16-64 host cores


def use_all_host_cores():
    """torchrun exports OMP_NUM_THREADS=1; the CPU arm is meant to use every core the process may run on."""
    from oracle import itk_oracle as orc

    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    orc.set_num_threads(n)
    return orc.num_threads()


def oracle_sample(size, fixed, moving, iters):
    """Bounded CPU sample: `iters` full-resolution Demons iterations of the oracle on all host cores."""
    from oracle import itk_oracle as orc

    cores = use_all_host_cores()
    p = orc.demons_params((1.5, 1.5, 1.5), iters, smooth_update_field=True)
    gf = orc.geom_of(fixed)
    t0 = time.perf_counter()
    _, st = orc.demons_execute(fixed.array, gf, moving.array, gf, p)
    dt = time.perf_counter() - t0
    vox_it = fixed.GetNumberOfPixels() * st["elapsed_iterations"]
    return vox_it / dt / 1e6, dt, st["elapsed_iterations"], cores


def oracle_whole_registration(fixed, moving):
    """The whole `fast_symmetric_forces_demons_registration` of the headline configuration on the CPU oracle."""
    from oracle import platipy_ref as ref

    cores = use_all_host_cores()
    stats = []
    t0 = time.perf_counter()
    ref.fast_symmetric_forces_demons_registration(fixed, moving, resolution_staging=RES_STAGING, iteration_staging=ITER_STAGING, level_stats=stats)
    dt = time.perf_counter() - t0
    vox_it = float(sum(s["voxels"] * s["elapsed_iterations"] for s in stats))
    return vox_it / dt / 1e6, dt, [s["elapsed_iterations"] for s in stats], cores


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  SimpleITK/ITK cannot be installed offline, so this arm
    times the oracle port (the C restatement of the ITK filters, OpenMP on all host cores).  The first timed step is the WHOLE
    registration of the headline configuration -- pyramid, three levels with the reference's early stop, level glue, final warp --
    and gives `value` and `ms_per_step`; warm-up steps and the remaining timed steps are bounded samples (2 full-resolution
    iterations each) so that the run stays within a few minutes whatever K + W is; their median is reported beside it."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from platipy_b200.synth import synth_pair

    size = tuple(args.size)
    fixed, moving = synth_pair(size, seed=0, moving_seed=100)
    cores = use_all_host_cores()
    for _ in range(args.warmup):
        oracle_sample(size, fixed, moving, 1)
    value, dt, elapsed, cores = oracle_whole_registration(fixed, moving)
    samples = []
    for _ in range(max(0, args.steps - 1)):
        if sum(d for _, d in samples) > 90.0:
            break
        v, d, _, _ = oracle_sample(size, fixed, moving, 2)
        samples.append((v, d))
    sample = (f"whole registration once ({size[0]}x{size[1]}x{size[2]}, pyramid {RES_STAGING}, elapsed iterations {elapsed}, {dt:.1f} s on {cores} threads: "
              f"pyramid + levels + glue + final warp) -> value; {len(samples)} further bounded samples of 2 full-resolution iterations"
              + (f", median {statistics.median(v for v, _ in samples):.1f} {UNIT}" if samples else "")
              + "; CPU restatement of the ITK filters, not SimpleITK")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dt, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(size, args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))
    return 0


def workload_config(size, n_gpus):
    return {"workload": f"single-pair Demons {size[0]}x{size[1]}x{size[2]}, 3-level pyramid {RES_STAGING}, {ITER_STAGING} iters "
                        f"(BASELINE.json configs[1])" + ("" if n_gpus == 1 else f"; one independent pair per GPU x{n_gpus}"),
            "size": list(size), "resolution_staging": RES_STAGING, "iteration_staging": ITER_STAGING, "field_dtype": "f64",
            "cache": "inputs (2 x 268 MB) and fields (1.6 GB each) exceed the 126 MB L2; no explicit flush",
            "parallelism": "1 pair/GPU" if n_gpus > 1 else "single GPU"}


# ---------------------------------------------------------------------------------------------------------------------
# fusion workloads (BASELINE.json configs[3], configs[4])
# ---------------------------------------------------------------------------------------------------------------------
FUSION_CONFIGS = {
    "cfg4": {"size": (256, 256, 160), "n_atlases": 4, "n_structures": 5, "fusion": "vote", "baseline_config": 3},
    "cfg5": {"size": (512, 512, 256), "n_atlases": 8, "n_structures": 20, "fusion": "staple", "baseline_config": 4},
}


def fusion_settings(mode):
    """Pipeline settings of examples/atlas_segmentation.ipynb cell 20 with Demons [4, 2, 1] x [50, 50, 25]."""
    return {
        "auto_crop_target_image_settings": {"expansion_mm": [20, 20, 40]},
        "linear_registration_settings": {"reg_method": "similarity", "shrink_factors": [8, 4, 2], "smooth_sigmas": [4, 2, 0], "sampling_rate": 1,
                                         "default_value": -1000, "number_of_iterations": 50, "metric": "mean_squares", "optimiser": "gradient_descent",
                                         "verbose": False},
        "deformable_registration_settings": {"isotropic_resample": False, "resolution_staging": [4, 2, 1], "iteration_staging": [50, 50, 25],
                                             "smoothing_sigmas": [4, 2, 0], "ncores": 32, "default_value": -1000, "verbose": False},
        "label_fusion_settings": {"vote_type": "unweighted", "vote_params": None, "optimal_threshold": {}, "fusion": mode},
        "postprocessing_settings": {"run_postprocessing": True, "binaryfillhole_mm": 3, "structures_for_binaryfillhole": [],
                                    "structures_for_overlap_correction": []},
    }


def run_fusion(name, eng, dist, rank, world, runs=2):
    import numpy as np
    import torch

    from platipy_b200 import multiatlas
    from platipy_b200.engine import DeviceImage
    from platipy_b200.synth import synth_atlas_case

    cfg = FUSION_CONFIGS[name]
    size, n_atlas, n_struct = cfg["size"], cfg["n_atlases"], cfg["n_structures"]
    spacing = (1.0, 1.0, 1.5)
    ident = (1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0)
    names = [f"S{k:02d}" for k in range(n_struct)]
    ids = [f"{a:03d}" for a in range(n_atlas)]

    def dimg(t, dt):
        return DeviceImage(t, dt, spacing, (0.0, 0.0, 0.0), ident, False)

    with torch.cuda.device(eng.device):
        ct, truth = synth_atlas_case(size, n_struct, spacing, seed=0, atlas_seed=None, as_tensors=True)
        target = dimg(ct, np.float32)
        mine = multiatlas.shard_atlases(ids, rank, world)
        atlas_part = {}
        for a in mine:
            act, alabs = synth_atlas_case(size, n_struct, spacing, seed=0, atlas_seed=int(a), as_tensors=True)
            entry = {"CT Image": dimg(act, np.float32)}
            entry.update({n: dimg(l, np.uint8) for n, l in zip(names, alabs)})
            atlas_part[a] = entry
    torch.cuda.synchronize()
    settings = fusion_settings(cfg["fusion"])

    def barrier():
        eng.synchronize()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def once(timings=None):
        return multiatlas.run_segmentation(target, atlas_part, settings, atlas_ids=ids, gather_probabilities=False, timings=timings)

    masks, _ = once()  # warm-up (allocator pools, NCCL channels, shared-memory opt-ins)
    walls = []
    l0 = eng.launch_count()
    for _ in range(runs):
        barrier()
        t0 = time.perf_counter()
        masks, probs = once()
        barrier()
        walls.append(1e3 * (time.perf_counter() - t0))
    launches = (eng.launch_count() - l0) / runs
    stages = {}
    barrier()
    once(stages)
    barrier()

    def allmax(x):
        if dist is None:
            return float(x)
        t = torch.tensor([x], dtype=torch.float64, device=eng.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    wall = min(allmax(w) for w in walls)
    stage_keys = ["setup_s", "linear_s", "deformable_s", "weight_map_s", "pack_s", "exchange_s", "finalise_s", "gather_s"]
    stages_ms = {k[:-2] + "_ms": 1e3 * allmax(stages.get(k, 0.0)) for k in stage_keys}
    # integer checksums of the final masks (every rank holds all of them) and Dice against the cohort's ground truth
    checks, dices = {}, []
    with torch.cuda.stream(eng.stream):
        idx = torch.arange(masks[names[0]].tensor.numel(), device=eng.device, dtype=torch.int64) % 1000003
        for n, tr in zip(names, truth):
            m = masks[n].tensor.reshape(-1).to(torch.int64)
            checks[n] = [int(m.sum().item()), int((m * idx).sum().item())]
            inter = int((m * tr.reshape(-1).to(torch.int64)).sum().item())
            dices.append(2.0 * inter / max(1, checks[n][0] + int(tr.sum().item())))
    return {"workload": f"run_segmentation {name}: {n_atlas} atlases x {n_struct} structures at {size[0]}x{size[1]}x{size[2]} (spacing 1 x 1 x 1.5 mm), "
                        f"similarity linear_registration -> Demons [4,2,1]x[50,50,25] -> label propagation -> "
                        f"{'STAPLE' if cfg['fusion'] == 'staple' else 'unweighted vote'} -> process_probability_image "
                        f"(BASELINE.json configs[{cfg['baseline_config']}])",
            "scaling": "strong", "n_gpus": world, "wall_ms": wall, "wall_ms_runs": [allmax(w) for w in walls],
            "atlases_per_rank_max": -(-n_atlas // world), "structures_per_rank_max": -(-n_struct // world),
            "stages_ms": stages_ms, "gpu_launches_per_run_rank0": launches,
            "exchange_payload_bytes_per_rank": int(stages.get("exchange_bytes", 0)), "cropped_grid": stages.get("grid"), "payload": stages.get("payload"),
            "mask_checksums": checks, "dice_vs_truth_min": min(dices), "dice_vs_truth_mean": sum(dices) / len(dices),
            "inputs": "resident in HBM on the owning rank", "outputs": "UInt8 masks of all structures on every rank (HBM); probabilities on the owner rank"}


def run_resample_cfg3(eng, size):
    """BASELINE.json configs[2]: one CT (linear, default -1000) + 20 UInt8 masks (nearest neighbour, 0) through a dense f64 DVF."""
    import torch

    from platipy_b200 import registration as reg
    from platipy_b200 import sitk_compat as sk
    from platipy_b200.sitk_compat import Image
    from platipy_b200.synth import smooth_random_dvf, synth_labels, synth_pair

    n = size[0] * size[1] * size[2]
    fixed, moving = synth_pair(size, seed=0, moving_seed=100)
    labels = [eng.to_device(Image(l)) for l in synth_labels(size, 20, seed=200)]
    dM, dF = eng.to_device(moving), eng.to_device(fixed)
    tfm = sk.DisplacementFieldTransform(Image(smooth_random_dvf(size, seed=9, peak_mm=6.0), is_vector=True))
    imgs = [dM] + labels
    dvs, ips = [-1000] + [0] * 20, [sk.sitkLinear] + [sk.sitkNearestNeighbor] * 20

    def timed(fn, reps=5):
        fn()
        fn()
        eng.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(eng.stream)
        for _ in range(reps):
            fn()
        e1.record(eng.stream)
        eng.synchronize()
        return e0.elapsed_time(e1) / reps

    def per_call():
        for im, dv, ip in zip(imgs, dvs, ips):
            reg.apply_transform(im, dF, tfm, dv, ip)

    ms_calls = timed(per_call)
    ms_batch = timed(lambda: reg.apply_transform_batch(imgs, dF, tfm, dvs, ips))
    peak, src = measured_peak_gbs()
    b_calls, b_batch = n * (32 + 20 * 26), n * (24 + 8 + 40)
    return {"workload": f"apply_transform {size[0]}x{size[1]}x{size[2]}: 1 CT (linear) + 20 UInt8 masks (nearest neighbour) through a dense f64 DVF "
                        "(BASELINE.json configs[2])",
            "per_call": {"ms": ms_calls, "algorithmic_bytes": b_calls, "achieved_gbs": b_calls / ms_calls / 1e6, "frac": b_calls / ms_calls / 1e6 / peak,
                         "bytes_per_voxel": "f32 linear 32, u8 NN 26 (DVF re-read per call)"},
            "batched": {"ms": ms_batch, "algorithmic_bytes": b_batch, "achieved_gbs": b_batch / ms_batch / 1e6, "frac": b_batch / ms_batch / 1e6 / peak,
                        "bytes_per_voxel": "24 DVF + 8 CT + 40 masks = 72", "Gvoxel_per_s": 21 * n / (ms_batch * 1e-3) / 1e9},
            "peak_gbs": peak, "peak_source": src}


def run_fast_mode(eng, fixed_p, moving_p, kw):
    """precision="fast" (float32 fields + float32 FMA smoothing inside the Demons loop; NOT a parity path) beside parity mode on the
    headline configuration: device time of the level loops and the error of the fast field against the parity field."""
    import torch

    from platipy_b200 import registration as reg

    dF, dM = eng.to_device(fixed_p), eng.to_device(moving_p)
    out, fields = {}, {}
    for mode in ("parity", "fast"):
        best = None
        for _ in range(3):
            eng.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(eng.stream)
            img, tfm, dvf = reg.fast_symmetric_forces_demons_registration(dF, dM, precision=mode, **kw)
            e1.record(eng.stream)
            eng.synchronize()
            st = reg.LAST_LEVEL_STATS
            total = e0.elapsed_time(e1)
            if best is None or total < best["registration_ms"]:
                best = {"registration_ms": total, "levels_ms": [s["gpu_ms"] for s in st], "elapsed_iterations": [s["elapsed_iterations"] for s in st],
                        "full_res_ms_per_iteration": st[-1]["gpu_ms"] / max(1, st[-1]["elapsed_iterations"])}
        out[mode] = best
        fields[mode] = dvf.tensor
    err = (fields["fast"] - fields["parity"]).abs().reshape(-1)
    sub = err[:: max(1, err.numel() // 16_000_000)][:16_000_000].to(torch.float64)
    q = torch.quantile(sub, torch.tensor([0.5, 0.99, 0.999], dtype=torch.float64, device=sub.device))
    out["error_vs_parity_mm"] = {"max": float(err.max()), "median": float(q[0]), "p99": float(q[1]), "p99.9": float(q[2]),
                                 "field_abs_max_mm": float(fields["parity"].abs().max()),
                                 "note": "fast mode is outside the 1e-4 mm parity bar (discontinuous thresholds of the ESM force amplify float32 rounding); "
                                         "it is an explicit switch, never the default and never the headline"}
    out["speedup_full_res_iteration"] = out["parity"]["full_res_ms_per_iteration"] / out["fast"]["full_res_ms_per_iteration"]
    out["speedup_registration"] = out["parity"]["registration_ms"] / out["fast"]["registration_ms"]
    return out


def run_default_staging(eng, fixed_p, moving_p):
    """The reference's own default arguments -- resolution_staging [8, 4, 1], 10 iterations per level (deformable.py:190-204) -- on the headline
    volume pair ("also report platipy's default staging"): device time of the whole registration, best of three."""
    import torch

    from platipy_b200 import registration as reg

    dF, dM = eng.to_device(fixed_p), eng.to_device(moving_p)
    kw = dict(resolution_staging=[8, 4, 1], iteration_staging=[10, 10, 10])
    best = None
    for _ in range(4):
        eng.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(eng.stream)
        reg.fast_symmetric_forces_demons_registration(dF, dM, **kw)
        e1.record(eng.stream)
        eng.synchronize()
        st = reg.LAST_LEVEL_STATS
        total = e0.elapsed_time(e1)
        if best is None or total < best["registration_ms"]:
            work = sum(s["voxels"] * s["elapsed_iterations"] for s in st)
            best = {"registration_ms": total, "levels_ms": [s["gpu_ms"] for s in st], "elapsed_iterations": [s["elapsed_iterations"] for s in st],
                    "glue_ms": total - sum(s["gpu_ms"] for s in st), "value": work / total / 1e3, "unit": UNIT}
    best.update(kw)
    return best


def run_b200(args):
    import numpy as np
    import torch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; platipy_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa_node(local) if world > 1 else None
    dist = None
    if world > 1:
        import torch.distributed as dist

        # rank 0 prints ONE JSON line on stdout: the image exports NCCL_DEBUG=VERSION, which makes NCCL print its version banner
        # there (NCCL_DEBUG_FILE does not move it).  An explicit WARN / INFO / TRACE choice of the caller is left alone.
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "NONE"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from platipy_b200 import registration as reg
    from platipy_b200.engine import Engine, pinned_image
    from platipy_b200.synth import synth_pair

    size = tuple(args.size)
    eng = Engine.get(local)
    fixed, moving = synth_pair(size, seed=0, moving_seed=100 + rank)
    fixed_p, moving_p = pinned_image(fixed), pinned_image(moving)
    dF, dM = eng.to_device(fixed_p), eng.to_device(moving_p)
    kw = dict(resolution_staging=RES_STAGING, iteration_staging=ITER_STAGING)
    nvox = fixed.GetNumberOfPixels()

    def barrier():
        eng.synchronize()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    level_stats = {}

    def device_step():
        img, tfm, dvf = reg.fast_symmetric_forces_demons_registration(dF, dM, **kw)
        level_stats["last"] = reg.LAST_LEVEL_STATS[:]
        return dvf

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = eng.launch_count()
        t0 = time.perf_counter()
        with torch.cuda.stream(eng.stream):
            e0.record(eng.stream)
            for _ in range(steps):
                out = fn()
                del out
            e1.record(eng.stream)
        barrier()
        wall = time.perf_counter() - t0
        return e0.elapsed_time(e1), wall, eng.launch_count() - l0

    # ---- device-resident arm ----
    for _ in range(args.warmup):
        device_step()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    dev_ms, dev_wall, launches = timed(device_step, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    stats = level_stats["last"]
    vox_it = float(sum(s["voxels"] * s["elapsed_iterations"] for s in stats))

    # ---- end-to-end arms (public host API, pinned host buffers in, host images out, copies inside the timed region) ----
    # e2e: K registrations through registration.iter_registrations, the package's form for back-to-back calls (the reference's use is a
    #      loop over atlases): every step uploads its own two inputs and downloads its own field + image; the copies of neighbouring
    #      steps overlap the compute.  e2e_single_call: K separate synchronous drop-in calls, nothing overlapped across calls.
    def consume(res):
        img, tfm, dvf = res
        return float(dvf.array[0, 0, 0, 0]) + float(img.array[0, 0, 0])

    def e2e_step():
        return consume(reg.fast_symmetric_forces_demons_registration(fixed_p, moving_p, **kw))

    def e2e_pipelined(steps):
        acc = 0.0
        for res in reg.iter_registrations(((fixed_p, moving_p) for _ in range(steps)), **kw):
            acc += consume(res)
            del res
        return acc

    for _ in range(max(1, min(args.warmup, 3))):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_single_ms = 1e3 * (time.perf_counter() - t0) / args.steps
    e2e_pipelined(4)  # warm-up: the pinned result buffers of the registrations in flight are allocated here (about a second per 1.6 GB)
    barrier()
    t0 = time.perf_counter()
    e2e_pipelined(args.steps)
    barrier()
    e2e_ms = 1e3 * (time.perf_counter() - t0) / args.steps

    def allmax(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=eng.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=eng.device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    ms_per_step = allmax(dev_ms / args.steps)
    e2e_ms = allmax(e2e_ms)
    e2e_single_ms = allmax(e2e_single_ms)
    total_vox_it = allsum(vox_it)
    total_launches = allsum(float(launches))
    value = total_vox_it / (ms_per_step * 1e-3) / 1e6
    e2e_value = total_vox_it / (e2e_ms * 1e-3) / 1e6
    del dF, dM

    # ---- the headline line is complete here; what follows only ADDS keys (fusion, resample_cfg3, fast_mode, cpu_baseline) --------
    line = None
    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        full = stats[-1]
        it_ms = full["gpu_ms"] / max(1, full["elapsed_iterations"])
        achieved = BYTES_PER_VOXEL_ITER_F64 * full["voxels"] / (it_ms * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get("demons_iteration_dram_bytes")
            except Exception:  # noqa: BLE001
                traffic = None
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                    "kernel": "full-resolution Demons iteration (warp + force + smooth-U + add/smooth-D kernels)",
                    "algorithmic_bytes_per_launch": BYTES_PER_VOXEL_ITER_F64 * full["voxels"], "ms_per_launch": it_ms, "peak_source": peak_src,
                    "iterations_per_s_fullres": 1e3 / it_ms}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(size, world),
                "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms, "steps": args.steps, "h2d_bytes_per_step": int(2 * nvox * 4 * world),
                        "d2h_bytes_per_step": int((nvox * 24 + nvox * 4) * world),
                        "api": "registration.iter_registrations: back-to-back registrations, each step's uploads / downloads overlap its neighbours' compute"},
                "e2e_single_call": {"value": total_vox_it / (e2e_single_ms * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": e2e_single_ms, "steps": args.steps,
                                    "api": "fast_symmetric_forces_demons_registration, one synchronous drop-in call per step (nothing overlaps across calls)"},
                "gpu_launches": int(total_launches), "roofline": roofline, "cpu_baseline": None, "clocks": clocks,
                "levels": [{"voxels": s["voxels"], "elapsed_iterations": s["elapsed_iterations"], "gpu_ms": s["gpu_ms"], "metric": s["metric"],
                            "rms_change": s["rms_change"]} for s in stats]}
        if numa is not None:
            line["numa_binding_rank0"] = numa

    # A rank that fails inside a collective of the extra workloads would leave the others waiting for ever and the measured line
    # unprinted: past this deadline rank 0 prints what it has and every rank leaves.
    printed = threading.Event()

    def emit():
        if rank == 0 and not printed.is_set():
            printed.set()
            print(json.dumps(line), flush=True)

    def deadline():
        if line is not None:
            line.setdefault("extras_note", f"the extra workloads did not finish within {args.extras_timeout} s; keys after the headline may be missing")
        emit()
        os._exit(0)

    watchdog = threading.Timer(float(args.extras_timeout), deadline)
    watchdog.daemon = True
    watchdog.start()

    # ---- the fusion workloads, cfg3 and fast mode (their own keys; they feed nothing above) ----
    if not args.no_fusion:
        fusion = {}
        for name in args.fusion_configs:
            try:
                torch.cuda.empty_cache()
                fusion[name] = run_fusion(name, eng, dist, rank, world)
            except Exception as e:  # noqa: BLE001 -- a failure here is reported, it must not cost the headline line
                fusion[name] = {"error": repr(e)[:400]}
            if line is not None:
                line["fusion"] = fusion
    if world == 1 and not args.no_cfg3:
        try:
            torch.cuda.empty_cache()
            line["resample_cfg3"] = run_resample_cfg3(eng, size)
        except Exception as e:  # noqa: BLE001
            line["resample_cfg3"] = {"error": repr(e)[:400]}
    if world == 1 and not args.no_cfg3:
        try:
            torch.cuda.empty_cache()
            line["platipy_default_staging"] = run_default_staging(eng, fixed_p, moving_p)
        except Exception as e:  # noqa: BLE001
            line["platipy_default_staging"] = {"error": repr(e)[:400]}
    if world == 1 and not args.no_fast_mode:
        try:
            torch.cuda.empty_cache()
            line["fast_mode"] = run_fast_mode(eng, fixed_p, moving_p, kw)
        except Exception as e:  # noqa: BLE001
            line["fast_mode"] = {"error": repr(e)[:400]}
    # bounded CPU sample on the host cores of this box (rank 0, N = 1 only)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, dt, it, cores = oracle_sample(size, fixed, moving, CPU_SAMPLE_ITERS)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"{it} full-resolution Demons iterations ({size[0]}x{size[1]}x{size[2]}) of the oracle port, {dt:.1f} s; "
                                          "CPU restatement of the ITK filters, not SimpleITK"}
    watchdog.cancel()
    emit()
    if dist is not None:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--size", type=int, nargs=3, default=[512, 512, 256])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-fusion", action="store_true", help="skip the run_segmentation workloads (cfg4, cfg5)")
    ap.add_argument("--fusion-configs", nargs="*", default=["cfg4", "cfg5"], choices=sorted(FUSION_CONFIGS))
    ap.add_argument("--no-cfg3", action="store_true", help="skip the apply_transform workload (cfg3)")
    ap.add_argument("--no-fast-mode", action="store_true", help="skip the fast-mode (float32 field) comparison")
    ap.add_argument("--extras-timeout", type=int, default=600, help="seconds after which the line is printed without the unfinished extra workloads")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
