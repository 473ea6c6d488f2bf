"""CPU (gloo, world_size 2): the N>1 host logic -- contiguous trajectory sharding and the
end-of-batch all-reduce of the pose-error scalars -- reproduces the single-process totals."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from captra_b200 import shard


def test_shard_range_covers_everything():
    for total, world in ((256, 8), (32, 1), (10, 4), (3, 8), (0, 2)):
        spans = [shard.shard_range(total, world, r) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
            assert a1 == b0 and a0 <= a1
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1
    assert shard.shard_range(256, 8, 3) == (96, 128)           # cfg4: 32 per GPU
    assert [shard.category_of(i) for i in range(8)] == [0, 1, 2, 3, 4, 5, 0, 1]


def test_cfg4_sharding_and_category_grouping():
    """BASELINE cfg4 host logic: 256 trajectories, category = global index mod 6, contiguous shards, grouped by category
    inside a rank.  Whatever the number of ranks, every trajectory is tracked exactly once with its own category."""
    from captra_b200 import track
    total = 256
    want = {}
    for i in range(total):
        want[track.NOCS_CATEGORIES[shard.category_of(i)]] = want.get(track.NOCS_CATEGORIES[shard.category_of(i)], 0) + 1
    for world in (1, 2, 4, 8, 3):
        seen, got = [], {}
        for rank in range(world):
            a, b = shard.shard_range(total, world, rank)
            ids = [shard.category_of(i) for i in range(a, b)]
            order, names = track.group_by_category(ids)
            assert sorted(order) == list(range(b - a))
            assert [track.NOCS_CATEGORIES[ids[i]] for i in order] == names           # the permutation groups, it does not relabel
            assert names == sorted(names, key=track.NOCS_CATEGORIES.index)           # contiguous spans, one per category
            seen += [a + i for i in order]
            for n in names:
                got[n] = got.get(n, 0) + 1
        assert sorted(seen) == list(range(total)) and got == want
    assert set(want) == set(track.NOCS_CATEGORIES) and all(track.CATEGORIES[c]["num_parts"] == 1 for c in want)
    # obj_info_nocs.yml: bottle, bowl, can are symmetric; camera, laptop, mug are not
    assert [track.CATEGORIES[c]["sym"] for c in track.NOCS_CATEGORIES] == [True, True, False, True, False, False]


def _poses(n, seed):
    g = torch.Generator().manual_seed(seed)
    mk = lambda: {"rotation": torch.randn(n, 2, 3, 3, generator=g), "translation": torch.randn(n, 2, 3, 1, generator=g),
                  "scale": torch.rand(n, 2, generator=g)}
    return mk(), mk()


def _worker(rank, world, port, total, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pose, gt = _poses(total, 0)                    # same global batch on every rank
    a, b = shard.shard_range(total, world, rank)
    sl = lambda d: {k: v[a:b] for k, v in d.items()}
    vec = shard.pose_error_scalars(sl(pose), sl(gt))
    shard.all_reduce_scalars(vec)
    if rank == 0:
        out.put(vec.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_all_reduce_matches_single_process():
    total, world = 37, 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get()
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    pose, gt = _poses(total, 0)
    want = shard.pose_error_scalars(pose, gt).numpy()
    np.testing.assert_allclose(got, want, rtol=1e-5)
    assert got[3] == total * 2
