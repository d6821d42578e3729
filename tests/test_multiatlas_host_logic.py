"""Host-side logic of the multi-GPU path on CPU: atlas sharding and the one exchange step, with a
world_size-2 gloo process group (no GPU; the kernels themselves are covered by the -m gpu tests)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from platipy_b200.multiatlas import atlas_bit, exchange_sum, shard_atlases


def test_shard_atlases_is_a_partition():
    ids = [f"{i:03d}" for i in (7, 1, 5, 3, 9, 2, 8)]
    for world in (1, 2, 3, 4, 8):
        shards = [shard_atlases(ids, r, world) for r in range(world)]
        flat = [a for s in shards for a in s]
        assert sorted(flat) == sorted(ids) and len(set(flat)) == len(ids)
        assert max(len(s) for s in shards) - min(len(s) for s in shards) <= 1
    assert shard_atlases(ids, 0, 2) == ["001", "003", "007", "009"]
    assert [atlas_bit(ids, a) for a in sorted(ids)] == list(range(7))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(0)  # every rank builds the same synthetic atlas set
        n_atlas, shape = 5, (6, 7, 8)
        ids = [f"{a:03d}" for a in range(n_atlas)]
        labels = {a: (rng.random(shape) > 0.5).astype(np.uint8) for a in ids}
        weights = {a: rng.random(shape).astype(np.float32) for a in ids}
        mine = shard_atlases(ids, rank, world)
        # weighted vote: local float32 sums in atlas order, then ONE all-reduce
        num = np.zeros(shape, np.float32)
        den = np.zeros(shape, np.float32)
        ones = np.zeros(shape, np.float32)
        packed = np.zeros(shape, np.int32)
        for a in mine:
            num = num + weights[a] * labels[a].astype(np.float32)
            den = den + weights[a]
            ones = ones + labels[a].astype(np.float32)
            packed |= (labels[a].astype(np.int32) << atlas_bit(ids, a))
        t = [torch.from_numpy(v) for v in (num, den, ones, packed)]
        exchange_sum(t)
        # references computed without sharding
        num_ref = np.zeros(shape, np.float32)
        den_ref = np.zeros(shape, np.float32)
        ones_ref = np.zeros(shape, np.float32)
        for a in ids:
            num_ref = num_ref + weights[a] * labels[a].astype(np.float32)
            den_ref = den_ref + weights[a]
            ones_ref = ones_ref + labels[a].astype(np.float32)
        ok = True
        ok &= np.allclose(t[0].numpy(), num_ref, rtol=1e-6, atol=1e-6)   # float32 sums, different association
        ok &= np.allclose(t[1].numpy(), den_ref, rtol=1e-6, atol=1e-6)
        ok &= np.array_equal(t[2].numpy(), ones_ref)                      # unweighted votes are exact
        for a in ids:                                                      # STAPLE: SUM of disjoint bits == OR
            ok &= np.array_equal((t[3].numpy() >> atlas_bit(ids, a)) & 1, labels[a])
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_vote_exchange_world_size_2_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res) == [(0, True), (1, True)]


def test_exchange_is_noop_without_process_group():
    t = torch.ones(4)
    assert exchange_sum([t])[0] is t and torch.equal(t, torch.ones(4))


# ---- structure-sharded fusion tail ---------------------------------------------------------------------------------------------------
def test_shard_structures_layout():
    from platipy_b200.multiatlas import mask_dtype, shard_structures

    names = [f"S{i:02d}" for i in range(20)]
    for world in (1, 2, 3, 4, 8):
        per, slots = shard_structures(names, world)
        assert per == -(-20 // world) and len(slots) == world and all(len(s) == per for s in slots)
        flat = [s for r in slots for s in r if s is not None]
        assert sorted(flat) == names
        assert max(sum(s is not None for s in r) for r in slots) - min(sum(s is not None for s in r) for r in slots) <= 1
    assert shard_structures(names, 8)[1][3] == ["S03", "S11", "S19"] and shard_structures(names, 8)[1][7] == ["S07", "S15", None]
    assert mask_dtype(8) == torch.uint8 and mask_dtype(9) == torch.int16 and mask_dtype(17) == torch.int32


def _exchange_worker(rank, world, port, q):
    from platipy_b200.multiatlas import exchange_all_gather, exchange_reduce_scatter, shard_structures

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(1)
        n_atlas, n_struct, shape = 11, 5, (3, 4, 5)  # 11 atlases: a 16-bit decision mask, reduced byte-wise
        dec = (rng.random((n_atlas, n_struct) + shape) > 0.5).astype(np.int16)
        per, slots = shard_structures(range(n_struct), world)
        mine = [a for a in range(n_atlas) if a % world == rank]
        stack = torch.zeros((world * per,) + shape, dtype=torch.int16)
        counts = torch.zeros((world * per,) + shape, dtype=torch.uint8)
        for r in range(world):
            for j, s in enumerate(slots[r]):
                for a in mine:
                    if s is not None:
                        stack[r * per + j] |= torch.from_numpy(dec[a, s] << a)
                        counts[r * per + j] += torch.from_numpy(dec[a, s].astype(np.uint8))
        block, cblock = exchange_reduce_scatter(stack), exchange_reduce_scatter(counts)
        ok = block.shape[0] == per
        for j, s in enumerate(slots[rank]):
            if s is None:
                ok &= not bool(block[j].any())
                continue
            expect = np.zeros(shape, np.int64)
            for a in range(n_atlas):
                expect |= dec[a, s].astype(np.int64) << a
            ok &= np.array_equal(block[j].numpy().astype(np.int64) & 0xFFFF, expect)            # SUM of disjoint bits == OR
            ok &= np.array_equal(cblock[j].numpy(), dec[:, s].sum(axis=0).astype(np.uint8))       # unweighted votes are exact
        gathered = exchange_all_gather(cblock)
        for r in range(world):
            for j, s in enumerate(slots[r]):
                if s is not None:
                    ok &= np.array_equal(gathered[r * per + j].numpy(), dec[:, s].sum(axis=0).astype(np.uint8))
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_reduce_scatter_and_all_gather_world_size_2_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_exchange_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res) == [(0, True), (1, True)]


def _pipeline_worker(rank, world, port, q):
    """run_segmentation, sharded over two gloo ranks, on the oracle-backed Engine stand-in (tests/fake_engine.py): atlas shards,
    the reduce-scatter over structures, the owner-only finalisation and the mask gather must reproduce the single-process result."""
    from platipy_b200 import multiatlas
    from platipy_b200.engine import Engine
    from platipy_b200.sitk_compat import Image
    from platipy_b200.synth import synth_labels, synth_pair
    from tests.fake_engine import FakeEngine

    eng = FakeEngine(None)
    Engine.get = classmethod(lambda cls, device=None: eng)
    size, sp = (24, 20, 12), (1.0, 1.0, 1.5)
    target, _ = synth_pair(size, seed=0, spacing=sp, peak_mm=2.0)
    base = synth_labels(size, 3, seed=500)
    atlas_set = {}
    for a in range(3):
        _, ct = synth_pair(size, seed=0, spacing=sp, peak_mm=2.0, moving_seed=100 + a)
        entry = {"CT Image": ct}
        for s in range(3):
            if not (a == 2 and s == 1):  # atlas 002 lacks structure S1: holder sets differ between structures
                entry[f"S{s}"] = Image(np.roll(base[s], a - 1, axis=2), sp)
        atlas_set[f"{a:03d}"] = entry

    def settings_for(mode, vote):
        return {"deformable_registration_settings": {"isotropic_resample": False, "resolution_staging": [2], "iteration_staging": [3], "default_value": -1000},
                "label_fusion_settings": {"vote_type": vote, "vote_params": {"factor": 1e6, "sigma": 2.0, "epsilon": 1e-5, "normalise": False},
                                          "optimal_threshold": {}, "fusion": mode},
                "postprocessing_settings": {"run_postprocessing": True, "binaryfillhole_mm": 1, "structures_for_binaryfillhole": ["S0"],
                                            "structures_for_overlap_correction": ["S0", "S2"]}}

    cases = [("vote", "unweighted"), ("staple", "unweighted"), ("vote", "global")]
    single = {c: multiatlas.run_segmentation(target, atlas_set, settings_for(*c)) for c in cases}
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ok = True
        for c in cases:
            # every rank hands over its own shard only, with the full id list (the other form: the full dictionary)
            mine = multiatlas.shard_atlases(sorted(atlas_set), rank, world)
            part = {a: atlas_set[a] for a in mine}
            res, prob = multiatlas.run_segmentation(target, part, settings_for(*c), atlas_ids=sorted(atlas_set))
            res_o, prob_o = multiatlas.run_segmentation(target, part, settings_for(*c), atlas_ids=sorted(atlas_set), gather_probabilities=False)
            ok &= sorted(res) == sorted(single[c][0]) == ["S0", "S1", "S2"] and sorted(prob) == ["S0", "S1", "S2"]
            ok &= sorted(prob_o) == [s for i, s in enumerate(["S0", "S1", "S2"]) if i % world == rank]  # owners only
            for s in res:
                ok &= np.array_equal(res[s].array, single[c][0][s].array) and np.array_equal(res_o[s].array, single[c][0][s].array)
                if c[1] == "global":   # float32 sums change association with the partition
                    ok &= np.allclose(prob[s].array, single[c][1][s].array, rtol=1e-5, atol=1e-6)
                else:                  # counts and decision bits are exact under any partition
                    ok &= np.array_equal(prob[s].array, single[c][1][s].array)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_sharded_run_segmentation_world_size_2_gloo_on_the_fake_engine(built):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_pipeline_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in range(world)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert sorted(res) == [(0, True), (1, True)]
