#!/usr/bin/env python
"""
bench.py -- headline benchmark of the Demons hot path (BASELINE.json configs[1]).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--size X Y Z]

A "step" is one `fast_symmetric_forces_demons_registration` of a synthetic 512x512x256 float32 pair,
3-level pyramid [4, 2, 1], 100/50/25 iterations (early stop as in the reference).  N > 1: one pair per rank
(rank r registers atlas 100+r to the common target: independent units, weak scaling) followed by the
path's one exchange step, the NCCL all-reduce of the propagated-label vote volume.

metric   demons_voxel_iterations_per_s, unit Mvoxel*it/s: (sum over levels of voxels x elapsed iterations,
         summed over ranks) / (max over ranks of the device time of the K timed steps / K).
value    inputs already resident in HBM (device-in, device-out call).
e2e      the same metric through the public host API: pinned host buffers in, host images out; the H2D
         and D2H copies are inside the timed region.
roofline the full-resolution Demons iteration (the kernels between two iterations of level 2), algorithmic
         176 B/voxel/iteration with f64 fields, against MEASURED_PEAKS.json hbm_gbs.
cpu_baseline / --impl reference: the CPU oracle (the restatement of the ITK filters; SimpleITK itself is
         not installable offline) timed on a bounded sample on all host cores.
experiments  (optional, N = 1, once per box, --no-experiments to skip) informational measurements taken in child processes AFTER
         every key above has been measured: A/B of compiled-out kernel variants, a pipelined end-to-end run, platipy's default
         staging, the newest GPU tests without -x.  Bounded to 200 s; feeds no other key.
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


CPU_SAMPLE_ITERS = 10  # full-resolution iterations per bounded CPU sample: 10-25 s on 16-64 host cores


def oracle_sample(size, fixed, moving, iters):
    """Bounded CPU sample: `iters` full-resolution Demons iterations of the oracle on all host cores."""
    from oracle import itk_oracle as orc

    p = orc.demons_params((1.5, 1.5, 1.5), iters, smooth_update_field=True)
    gf = orc.geom_of(fixed)
    t0 = time.perf_counter()
    _, st = orc.demons_execute(fixed.array, gf, moving.array, gf, p)
    dt = time.perf_counter() - t0
    vox_it = fixed.GetNumberOfPixels() * st["elapsed_iterations"]
    return vox_it / dt / 1e6, dt, st["elapsed_iterations"], orc.num_threads()


def run_experiments(timeout_s=100, budget_s=200):
    """A/B of kernel variants that are compiled out of the default library, in CHILD processes (their own CUDA contexts) after
    every number of the JSON line has been measured: the TMA staging forms of the fused smoothing kernel (row-wise bulk copies,
    one tensor-map copy per plane tile) against cp.async, via profiles/ab_variants.py, which also reports whether the displacement
    field is bit-identical.  Only runs when the alternative build platipy_b200/libb200reg_tma.so is present
    (make -C platipy_b200/csrc OUT=../libb200reg_tma.so EXTRA="-DB200REG_ENABLE_ZM_TMA -DB200REG_AB_VARIANTS"); bounded by `timeout_s`; any failure is
    recorded as text and never touches the measured values.  Informational: not part of metric / value / e2e / roofline."""
    import subprocess

    import tempfile

    t_start = time.perf_counter()
    lib = "libb200reg_tma.so"
    if not os.path.exists(os.path.join(ROOT, "platipy_b200", lib)):
        return None  # the alternative build is the switch for the whole informational block
    # once per box: a scaling run repeats the N = 1 command, the informational block need not run again
    marker = os.path.join(tempfile.gettempdir(), "b200reg_bench_experiments_done")
    if os.path.exists(marker) and os.environ.get("B200REG_BENCH_EXPERIMENTS") != "force":
        return {"skipped": "already ran on this box (" + marker + ")"}
    try:
        open(marker, "w").write(str(time.time()))
    except OSError:
        pass
    # most informative first (the cap may cut the list short); the row-wise bulk copies were already measured slower in round 1
    specs = ["default=", f"tma_tensor=B200REG_ZM_TMA=2,lib={lib}", f"tma_tensor_tx64=B200REG_ZM_TMA=2,B200REG_ZM_TX32=0,lib={lib}",
             f"cp_async=B200REG_ZM_TMA=0,lib={lib}", f"cp_async_tx64=B200REG_ZM_TMA=0,B200REG_ZM_TX32=0,lib={lib}",
             f"tma_tensor_l2_128=B200REG_ZM_TMA=2,B200REG_ZM_TMA_L2=2,lib={lib}"]
    out = ""
    try:
        # own session: on a timeout the whole process group goes (the harness runs every configuration in a grandchild)
        proc = subprocess.Popen([sys.executable, os.path.join(ROOT, "profiles", "ab_variants.py")] + specs, stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                                text=True, start_new_session=True)
        try:
            out, err = proc.communicate(timeout=timeout_s)
            note = err[-300:] if proc.returncode else None
        except subprocess.TimeoutExpired:
            import signal

            os.killpg(proc.pid, signal.SIGKILL)
            out, err = proc.communicate()
            note = f"stopped after {timeout_s} s"
    except Exception as e:  # noqa: BLE001
        return {"error": repr(e)[:300]}
    res = {}
    for l in out.splitlines():  # one line per finished configuration: "<name> {json}"
        name, _, rest = l.partition(" ")
        if rest.startswith("{") and name in [sp.split("=")[0] for sp in specs]:
            try:
                res[name] = json.loads(rest)
            except ValueError:
                pass
        if l.startswith("AB {"):  # the harness's closing summary also carries the configurations that failed, with the error text
            try:
                for k, v in json.loads(l[3:]).items():
                    res.setdefault(k, v)
            except ValueError:
                pass
    for v in res.values():  # keep the line small
        if isinstance(v, dict) and "error" in v:
            v["error"] = str(v["error"])[-240:]
    exp = {"smoothing_tile_staging_ab": res, "what": "full-resolution ms per iteration and DVF identity of the fused smoothing kernel's staging variants "
                                                      "(profiles/ab_variants.py, child processes, after the timed regions)"}
    if note:
        exp["note"] = note
    # further child scripts, each printing one "EXP {json}" line:
    #   pipelined_e2e            copies of one registration overlapping the compute of its neighbours
    #   platipy_default_staging  the headline volume with platipy's own defaults ([8, 4, 1] shrink factors, 10 iterations per level), SURVEY 8d
    #   session3_rows            device time at the headline size of the entry points added without GPU time in round 1's third session
    #   session3_gpu_tests       the GPU tests of those entry points, run without -x (every outcome is recorded)
    # The whole block stays inside `budget_s`: a child gets what is left of it, and is skipped when that is under 15 s.
    for key, script, cap in (("pipelined_e2e", "exp_pipelined_e2e.py", 60), ("session3_gpu_tests", "exp_session3_tests.py", 90),
                             ("platipy_default_staging", "exp_default_staging.py", 45), ("session3_rows", "exp_session3_rows.py", 60)):
        cap = int(min(cap, budget_s - (time.perf_counter() - t_start)))
        if cap < 15:
            exp[key] = {"skipped": "time budget of the informational block used up"}
            continue
        try:
            proc = subprocess.Popen([sys.executable, os.path.join(ROOT, "profiles", script)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                                    start_new_session=True)
            try:
                out, err = proc.communicate(timeout=cap)
                got = [l for l in out.splitlines() if l.startswith("EXP ")]
                exp[key] = json.loads(got[-1][4:]) if got else {"error": (err or out)[-240:]}
            except subprocess.TimeoutExpired:
                import signal

                os.killpg(proc.pid, signal.SIGKILL)
                proc.communicate()
                exp[key] = {"error": f"stopped after {cap} s"}
        except Exception as e:  # noqa: BLE001
            exp[key] = {"error": repr(e)[:240]}
    return exp


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  SimpleITK/ITK cannot be installed
    offline, so this arm times the oracle port (the C restatement of the ITK filters, OpenMP on all host
    cores); each step is a bounded sample: CPU_SAMPLE_ITERS full-resolution Demons iterations."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from platipy_b200.synth import synth_pair

    size = tuple(args.size)
    fixed, moving = synth_pair(size, seed=0, moving_seed=100)
    vals = []
    cores = None
    # bounded: the whole --steps K --warmup W run stays within a few minutes whatever K + W is
    iters = max(2, min(CPU_SAMPLE_ITERS, 60 // max(1, args.warmup + args.steps)))
    for s in range(args.warmup + args.steps):
        v, dt, it, cores = oracle_sample(size, fixed, moving, iters)
        if s >= args.warmup:
            vals.append((v, dt))
    value = sum(v for v, _ in vals) / len(vals)
    ms = 1e3 * sum(dt for _, dt in vals) / len(vals)
    sample = f"{iters} full-resolution Demons iterations ({size[0]}x{size[1]}x{size[2]}) of the oracle port per step"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(size, args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))
    return 0


def workload_config(size, n_gpus):
    return {"workload": f"single-pair Demons {size[0]}x{size[1]}x{size[2]}, 3-level pyramid {RES_STAGING}, {ITER_STAGING} iters "
                        f"(BASELINE.json configs[1])" + ("" if n_gpus == 1 else f"; one pair per GPU x{n_gpus} + label vote all-reduce"),
            "size": list(size), "resolution_staging": RES_STAGING, "iteration_staging": ITER_STAGING, "field_dtype": "f64",
            "cache": "inputs (2 x 268 MB) and fields (1.6 GB each) exceed the 126 MB L2; no explicit flush",
            "parallelism": "1 pair/GPU" if n_gpus > 1 else "single GPU"}


def run_b200(args):
    import numpy as np
    import torch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; platipy_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        # rank 0 prints ONE JSON line on stdout: the image exports NCCL_DEBUG=VERSION, which makes NCCL print its version banner
        # there (NCCL_DEBUG_FILE does not move it).  An explicit WARN / INFO / TRACE choice of the caller is left alone.
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "NONE"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from platipy_b200 import registration as reg
    from platipy_b200 import sitk_compat as sk
    from platipy_b200.engine import Engine, pinned_image
    from platipy_b200.synth import synth_labels, synth_pair

    size = tuple(args.size)
    eng = Engine.get(local)
    fixed, moving = synth_pair(size, seed=0, moving_seed=100 + rank)
    fixed_p, moving_p = pinned_image(fixed), pinned_image(moving)
    dF, dM = eng.to_device(fixed_p), eng.to_device(moving_p)
    label = eng.to_device(sk.Image(synth_labels(size, 1, seed=200)[0])) if world > 1 else None
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
        level_stats["last"] = tfm_stats()
        if dist is not None:
            # the path's one exchange step: propagate a label, accumulate the vote, all-reduce, finalise
            lab = reg.apply_transform(label, dF, tfm, 0, sk.sitkNearestNeighbor)
            w = eng.weight_map(dF, img, 0)
            num = eng.empty(lab.tensor.shape, np.float32)
            den = eng.empty(lab.tensor.shape, np.float32)
            eng.vote_accumulate(lab, w, num, den, True)
            with torch.cuda.stream(eng.stream):
                dist.all_reduce(num)
                dist.all_reduce(den)
            eng.vote_finalize(num, den, dF, 1.0, 1e-4)
        return dvf

    def tfm_stats():
        return reg.LAST_LEVEL_STATS[:]

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

    # ---- end-to-end arm (public host API, pinned host buffers, copies inside the timed region) ----
    def e2e_step():
        img, tfm, dvf = reg.fast_symmetric_forces_demons_registration(fixed_p, moving_p, **kw)
        return float(dvf.array[0, 0, 0, 0]) + float(img.array[0, 0, 0])

    for _ in range(max(1, min(args.warmup, 2))):
        e2e_step()
    e2e_steps = max(1, min(args.steps, 3))
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_ms = 1e3 * (time.perf_counter() - t0) / e2e_steps

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
    total_vox_it = allsum(vox_it)
    total_launches = allsum(float(launches))
    value = total_vox_it / (ms_per_step * 1e-3) / 1e6
    e2e_value = total_vox_it / (e2e_ms * 1e-3) / 1e6

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
        # bounded CPU sample on the host cores of this box (rank 0, N = 1 only)
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            v, dt, it, cores = oracle_sample(size, fixed, moving, CPU_SAMPLE_ITERS)
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"{it} full-resolution Demons iterations ({size[0]}x{size[1]}x{size[2]}) of the oracle port, {dt:.1f} s; "
                             "CPU restatement of the ITK filters, not SimpleITK"}
        experiments = None
        if world == 1 and not args.no_experiments and os.environ.get("B200REG_BENCH_EXPERIMENTS", "1") != "0":
            try:
                experiments = run_experiments()
            except BaseException as e:  # noqa: BLE001 -- nothing here may cost the measured line
                experiments = {"error": repr(e)[:300]}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(size, world),
                "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms, "h2d_bytes_per_step": int(2 * nvox * 4 * world),
                        "d2h_bytes_per_step": int((nvox * 24 + nvox * 4) * world)},
                "gpu_launches": int(total_launches), "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
                "levels": [{"voxels": s["voxels"], "elapsed_iterations": s["elapsed_iterations"], "gpu_ms": s["gpu_ms"], "metric": s["metric"],
                            "rms_change": s["rms_change"]} for s in stats]}
        if experiments is not None:
            line["experiments"] = experiments
        print(json.dumps(line))
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
    ap.add_argument("--no-experiments", action="store_true", help="skip the informational A/B of compiled-out kernel variants")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
