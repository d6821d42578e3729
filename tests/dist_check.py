"""2-rank check of the sharded pipeline (run under torchrun on 2 GPUs):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_check.py
Every rank runs run_segmentation over its shard; the fused probabilities must equal the single-process run
(bit-exact for unweighted votes and for the STAPLE decision exchange)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from platipy_b200 import multiatlas
from tests.test_gpu_multiatlas import make_case

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
target, atlas_set = make_case(n_atlas=4)
ok = True
MODES = ("vote", "staple")


def settings_for(mode):
    return {"deformable_registration_settings": {"isotropic_resample": False, "resolution_staging": [2, 1], "iteration_staging": [8, 4]},
            "label_fusion_settings": {"vote_type": "unweighted", "vote_params": None, "optimal_threshold": {}, "fusion": mode}}


# no process group yet: every rank processes all atlases on its own GPU
single = {m: multiatlas.run_segmentation(target, atlas_set, settings_for(m))[1] for m in MODES}
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
for mode in MODES:
    _, sharded = multiatlas.run_segmentation(target, atlas_set, settings_for(mode))
    for s in single[mode]:
        same = np.array_equal(single[mode][s].array, sharded[s].array)
        print(f"rank {rank} {mode} {s}: sharded == single-process: {same}", flush=True)
        ok &= same
dist.barrier()
dist.destroy_process_group()
assert ok
print(f"rank {rank}: OK", flush=True)
