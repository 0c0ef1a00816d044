"""N > 1 host logic on CPU: world_size-2 gloo processes shard a clip, exchange only per-picture SADs (host side),
and must arrive at the single-process boundaries."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from video_transformer_b200 import scene, shard


def test_assign_videos_balances_longest_first():
    sizes = [1800] * 60 + [9000, 7200, 3600, 600]
    for world in (1, 2, 4, 8):
        a = shard.assign_videos(sizes, world)
        assert sorted(i for r in a for i in r) == list(range(len(sizes)))
        loads = [sum(sizes[i] for i in r) for r in a]
        assert max(loads) - min(loads) <= max(sizes)


def test_split_gop_aligned():
    kf = list(range(0, 3000, 30)) + [1234]
    for world in (1, 2, 4, 8):
        parts = shard.split_gop_aligned(kf, 0, 3000, world)
        assert parts[0][0] == 0 and parts[-1][1] == 3000
        for (a, b), (c, d) in zip(parts, parts[1:]):
            assert b == c and (c in kf)
    assert shard.split_gop_aligned([0], 0, 100, 4) == [(0, 0), (0, 0), (0, 0), (0, 100)]   # fewer GOPs than ranks


def _fake_sad(n, seed=3):
    rng = np.random.default_rng(seed)
    s = rng.integers(0, 1280 * 720 * 3, n).astype(np.uint64)
    s[[97, 400, 401, 777]] = 1280 * 720 * 90
    return s


def _worker(rank, world, port, n, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    kf = list(range(0, n, 30))
    a, b = shard.split_gop_aligned(kf, 0, n, world)[rank]
    full = _fake_sad(n)                     # stands for what this rank's GPU pass would have produced for [a,b)
    mine = (a, full[a:b])
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)   # 8 bytes per picture: host-side, not on the data path
    sad, scores, cuts = shard.merge_and_score(gathered, 1280, 720, 0.10)
    np.save(os.path.join(out_dir, "cuts_%d.npy" % rank), cuts)
    dist.destroy_process_group()


def test_two_rank_gloo_boundaries_equal_single_process(tmp_path):
    n = 1000
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, n, str(tmp_path)), nprocs=2, join=True)
    full = _fake_sad(n)
    full[0] = 0
    exp = scene.select_cuts(scene.scene_scores(full, 1280, 720), 0.10)
    for r in range(2):
        assert np.array_equal(np.load(tmp_path / ("cuts_%d.npy" % r)), exp)
    assert set([97, 400, 777]) <= set(exp.tolist())
