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
