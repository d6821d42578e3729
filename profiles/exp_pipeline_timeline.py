"""Timeline of the pipelined host API (registration.iter_registrations): CUDA events around every upload, registration and download
of four back-to-back registrations at the headline size, printed relative to the first upload.  Diagnostic for the e2e arm."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from platipy_b200 import registration as reg
from platipy_b200.engine import Engine, pinned_image
from platipy_b200.synth import synth_pair

eng = Engine.get(0)
fixed, moving = synth_pair((512, 512, 256), seed=0, moving_seed=100)
fp, mp = pinned_image(fixed), pinned_image(moving)
kw = dict(resolution_staging=[4, 2, 1], iteration_staging=[100, 50, 25])
marks = []
orig_async, orig_up = eng.to_host_async, eng.to_device_async


def mark(stream, name):
    e = torch.cuda.Event(enable_timing=True)
    e.record(stream)
    marks.append((name, e, time.perf_counter()))


def to_host_async(d):
    mark(eng.stream, "compute_tail_before_d2h")
    mark(eng.copy_out, "d2h_queue_head")
    out = orig_async(d)
    mark(eng.copy_out, f"d2h_done_{'field' if d.is_vector else 'image'}")
    return out


def to_device_async(im):
    mark(eng.copy_in, "h2d_start")
    out = orig_up(im)
    mark(eng.copy_in, "h2d_done")
    return out


eng.to_host_async, eng.to_device_async = to_host_async, to_device_async
for _ in reg.iter_registrations([(fp, mp)] * 2, **kw):
    pass
torch.cuda.synchronize()
marks.clear()
t0 = time.perf_counter()
mark(eng.stream, "t0")
n = 0
for res in reg.iter_registrations([(fp, mp)] * 4, **kw):
    mark(eng.stream, f"result_{n}_returned_to_caller")
    n += 1
    del res
torch.cuda.synchronize()
wall = time.perf_counter() - t0
base = marks[0][1]
rows = [(name, round(base.elapsed_time(e), 2), round(1e3 * (th - t0), 2)) for name, e, th in marks]
print("TIMELINE " + json.dumps({"wall_ms_per_step": 1e3 * wall / 4, "events (name, device ms, host ms at enqueue)": rows}))
