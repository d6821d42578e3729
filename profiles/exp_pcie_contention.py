"""What the host side of a GPU box gives N ranks at once: every rank copies the byte counts of one end-to-end Demons step (537 MB up,
1.88 GB down) between pinned host memory and its GPU, all ranks together, and reports GB/s per rank and summed.  Run with
torchrun at N = 1 and N = 8; the ratio is the ceiling of the end-to-end scaling efficiency of `bench.py`'s `e2e` (a platform
property: the copies already overlap the compute).  Prints: PCIE {json} on rank 0."""
import json
import os

import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", 0))
world = int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
UP, DOWN = 537_000_000, 1_880_000_000
h_up = torch.empty(UP, dtype=torch.uint8, pin_memory=True)
h_dn = torch.empty(DOWN, dtype=torch.uint8, pin_memory=True)
d_up = torch.empty(UP, dtype=torch.uint8, device="cuda")
d_dn = torch.empty(DOWN, dtype=torch.uint8, device="cuda")
s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()


def timed(fn, reps=4):
    fn()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    s_in.synchronize()
    s_out.synchronize()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    t = torch.tensor([ms], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def up():
    s_in.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s_in):
        d_up.copy_(h_up, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s_in)


def down():
    s_out.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s_out):
        h_dn.copy_(d_dn, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s_out)


def both():
    s_in.wait_stream(torch.cuda.current_stream())
    s_out.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s_in):
        d_up.copy_(h_up, non_blocking=True)
    with torch.cuda.stream(s_out):
        h_dn.copy_(d_dn, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s_in)
    torch.cuda.current_stream().wait_stream(s_out)


out = {"world": world}
for name, fn, nbytes in (("h2d", up, UP), ("d2h", down, DOWN), ("both", both, UP + DOWN)):
    ms = timed(fn)
    out[name] = {"ms_max_over_ranks": ms, "gbs_per_rank": nbytes / ms / 1e6, "gbs_all_ranks": world * nbytes / ms / 1e6}
if rank == 0:
    print("PCIE " + json.dumps(out))
if world > 1:
    dist.destroy_process_group()
