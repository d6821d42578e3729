"""N-rank check of the sharded pipeline (run under torchrun on 2, 4 or 8 GPUs):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_check.py
Every rank runs run_segmentation over its atlas shard and finalises its structure shard; masks and fused probabilities must equal
the single-process run -- bit-exact for unweighted votes (UInt8 counts) and for STAPLE (decision masks), to the float32
association tolerance for weighted votes.  Covers the full-dictionary and the own-shard-plus-id-list calling forms, a structure
that one atlas lacks, more ranks than structures / atlases (8 ranks), and post-processing on the owner rank."""
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
target, atlas_set = make_case(n_atlas=5, n_struct=3)
del atlas_set["004"]["S1"]  # holder sets differ between structures
ok = True
CASES = (("vote", "unweighted"), ("staple", "unweighted"), ("vote", "global"))


def settings_for(mode, vote):
    return {"deformable_registration_settings": {"isotropic_resample": False, "resolution_staging": [2, 1], "iteration_staging": [8, 4]},
            "label_fusion_settings": {"vote_type": vote, "vote_params": {"factor": 1e6, "sigma": 2.0, "epsilon": 1e-5, "normalise": False},
                                      "optimal_threshold": {}, "fusion": mode},
            "postprocessing_settings": {"run_postprocessing": True, "binaryfillhole_mm": 1, "structures_for_binaryfillhole": ["S0"],
                                        "structures_for_overlap_correction": ["S0", "S2"]}}


# no process group yet: every rank processes all atlases on its own GPU
single = {c: multiatlas.run_segmentation(target, atlas_set, settings_for(*c)) for c in CASES}
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ids = sorted(atlas_set)
for c in CASES:
    for form in ("full", "shard"):
        part = atlas_set if form == "full" else {a: atlas_set[a] for a in multiatlas.shard_atlases(ids, rank, world)}
        masks, probs = multiatlas.run_segmentation(target, part, settings_for(*c), atlas_ids=ids)
        for s in single[c][0]:
            same_m = np.array_equal(single[c][0][s].array, masks[s].array)
            if c[1] == "unweighted":
                same_p = np.array_equal(single[c][1][s].array, probs[s].array)
            else:
                same_p = np.allclose(single[c][1][s].array, probs[s].array, rtol=1e-5, atol=1e-6)
            print(f"rank {rank} {c} {form} {s}: masks equal {same_m}, probabilities equal {same_p}", flush=True)
            ok &= same_m and same_p
dist.barrier()
dist.destroy_process_group()
assert ok
print(f"rank {rank}: OK", flush=True)
