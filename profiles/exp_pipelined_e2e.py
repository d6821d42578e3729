"""Experiment (informational): end-to-end throughput of back-to-back registrations when the host<->device copies of one
registration overlap the compute of its neighbours.  The public host API is synchronous like the reference's (inputs up, compute,
results down: 172 ms per 512x512x256 registration of which 42 ms are PCIe time); an atlas pipeline registers N pairs in a row,
so the copies can ride a second stream:

    copy stream   : H2D pair k+1 ........ | ...... D2H results k-1 (DVF as [z, y, x, 3] float64 + registered image)
    engine stream :        compute pair k (device-in / device-out call, unchanged kernels)

Everything a step needs still crosses PCIe inside the timed region (pinned host buffers both ways).  Prints one JSON line.
Uses torch only for the copies / the layout change of the result (plumbing)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from platipy_b200 import registration as reg
from platipy_b200.engine import DeviceImage, Engine, pinned_image
from platipy_b200.synth import synth_pair

SIZE = (512, 512, 256)
STEPS = int(os.environ.get("EXP_STEPS", "6"))
eng = Engine.get(0)
dev = eng.device
fixed, moving = synth_pair(SIZE, seed=0, moving_seed=100)
fixed_p, moving_p = pinned_image(fixed), pinned_image(moving)
kw = dict(resolution_staging=[4, 2, 1], iteration_staging=[100, 50, 25])
z, y, x = fixed.array.shape
copy = torch.cuda.Stream(device=dev)
host_f, host_m = torch.from_numpy(fixed_p.array), torch.from_numpy(moving_p.array)
# two sets of device inputs and pinned outputs (double buffering)
d_in = [(torch.empty((z, y, x), dtype=torch.float32, device=dev), torch.empty((z, y, x), dtype=torch.float32, device=dev)) for _ in range(2)]
h_out = [(torch.empty((z, y, x, 3), dtype=torch.float64, pin_memory=True), torch.empty((z, y, x), dtype=torch.float32, pin_memory=True)) for _ in range(2)]
up_done = [torch.cuda.Event() for _ in range(2)]
comp_done = [torch.cuda.Event() for _ in range(2)]
down_done = [torch.cuda.Event() for _ in range(2)]


def upload(k):
    b = k & 1
    with torch.cuda.stream(copy):
        copy.wait_event(comp_done[b])  # the compute that last read this input buffer has finished (no-op the first time round)
        d_in[b][0].copy_(host_f, non_blocking=True)
        d_in[b][1].copy_(host_m, non_blocking=True)
        up_done[b].record(copy)


def compute(k):
    b = k & 1
    cur = torch.cuda.current_stream(dev)
    cur.wait_event(up_done[b])
    f = DeviceImage(d_in[b][0], np.float32, fixed.GetSpacing(), fixed.GetOrigin(), fixed.GetDirection())
    m = DeviceImage(d_in[b][1], np.float32, moving.GetSpacing(), moving.GetOrigin(), moving.GetDirection())
    img, tfm, dvf = reg.fast_symmetric_forces_demons_registration(f, m, **kw)  # device in -> device out, no host synchronisation
    comp_done[b].record(cur)
    return img, dvf


def download(k, img, dvf):
    b = k & 1
    with torch.cuda.stream(copy):
        copy.wait_event(comp_done[b])
        copy.wait_event(down_done[b])  # the previous download into this pinned buffer is complete
        h_out[b][0].copy_(dvf.tensor.permute(1, 2, 3, 0), non_blocking=True)  # SoA planes -> [z, y, x, 3] on the way out
        h_out[b][1].copy_(img.tensor, non_blocking=True)
        down_done[b].record(copy)
        img.tensor.record_stream(copy)
        dvf.tensor.record_stream(copy)


def run(steps):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    upload(0)
    for k in range(steps):
        if k + 1 < steps:
            upload(k + 1)
        img, dvf = compute(k)
        download(k, img, dvf)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / steps * 1e3


def sequential(steps):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        img, tfm, dvf = reg.fast_symmetric_forces_demons_registration(fixed_p, moving_p, **kw)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / steps * 1e3


run(2)
seq_ms = sequential(2)
seq_ms = sequential(3)
pipe_ms = run(STEPS)
stats = reg.LAST_LEVEL_STATS
vox_it = float(sum(s["voxels"] * s["elapsed_iterations"] for s in stats))
ref_dvf = eng.to_host(compute(0)[1], pinned=False).array
ok = bool(np.array_equal(h_out[(STEPS - 1) & 1][0].numpy(), ref_dvf))
print("EXP " + json.dumps({"sequential_host_api_ms_per_registration": seq_ms, "pipelined_ms_per_registration": pipe_ms, "steps": STEPS,
                           "pipelined_Mvoxel_it_per_s": vox_it / (pipe_ms * 1e-3) / 1e6, "h2d_bytes_per_registration": int(2 * x * y * z * 4),
                           "d2h_bytes_per_registration": int(x * y * z * 28), "results_identical_to_synchronous_call": ok}))
