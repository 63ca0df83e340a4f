"""Multi-rank host logic on CPU: world_size-2 gloo process group (frames are sharded per rank with no data-path
collective; only timings / counts are reduced)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fv2p_b200 import sharding, synth
from oracle import oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_shard_range_partitions_every_frame_once():
    for n in (0, 1, 7, 8, 64, 65):
        for w in (1, 2, 4, 8):
            got = []
            for r in range(w):
                lo, hi = sharding.shard_range(n, r, w)
                assert 0 <= lo <= hi <= n and hi - lo in (n // w, n // w + 1)
                got += list(range(lo, hi))
            assert got == list(range(n))
    with pytest.raises(ValueError):
        sharding.shard_range(8, 2, 2)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        frames = [synth.lidar_frame("kitti", seed=s, az_steps=24) for s in range(5)]
        mine = sharding.shard_frames(frames)
        # each rank voxelizes its own frames with local batch ids (the CPU oracle stands in for the GPU path here:
        # this test is about the partition and the reductions, the CUDA parity is tests/test_gpu_parity.py)
        cfg = synth.DATASETS["kitti"]
        coords = [O.voxelize(f, cfg["voxel_size"], cfg["point_cloud_range"], 5, 4000)[1] for f in mine]
        local = O.collate(coords)
        assert local.shape[0] == sum(c.shape[0] for c in coords)
        assert (local[:, 0].max() if local.size else -1) == len(mine) - 1
        elapsed = sharding.max_over_ranks(1.0 + rank)
        counts = sharding.gather_counts([len(mine), local.shape[0]])
        q.put((rank, len(mine), int(local.shape[0]), elapsed, counts))
    finally:
        dist.destroy_process_group()


def test_two_ranks_gloo_shard_and_reduce():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[1] for r in res] == [3, 2]                       # 5 frames over 2 ranks
    assert all(r[3] == 2.0 for r in res)                       # MAX over ranks of (1+rank)
    assert res[0][4] == res[1][4] == [[3, res[0][2]], [2, res[1][2]]]
    # the union of the shards is the whole batch: voxel totals match a single-process run
    cfg = synth.DATASETS["kitti"]
    total = sum(O.voxelize(synth.lidar_frame("kitti", seed=s, az_steps=24), cfg["voxel_size"],
                           cfg["point_cloud_range"], 5, 4000)[1].shape[0] for s in range(5))
    assert res[0][2] + res[1][2] == total


def test_single_process_identities():
    assert sharding.max_over_ranks(3.5) == 3.5
    assert sharding.gather_counts([1, 2]) == [[1, 2]]
    assert len(sharding.shard_frames(list(range(6)), 1, 3)) == 2


def test_bench_shards_the_fixed_batch_into_equal_launches(monkeypatch):
    """bench.py's waymo_64 workload: the 64 frames are split per frame over the ranks and run in 4-frame launches;
    every frame is processed exactly once whatever the world size (frames stand in as their sequence numbers)."""
    import bench
    monkeypatch.setattr(bench, "make_frames", lambda wl, first, n: list(range(first, first + n)))
    wl = bench.WORKLOADS["waymo_64"]
    for world in (1, 2, 4, 8, 3):
        seen = []
        for rank in range(world):
            batches, (lo, hi) = bench._batches_for_rank(wl, rank, world)
            assert all(1 <= len(b) <= wl["batch"] for b in batches)
            flat = [f for b in batches for f in b]
            assert flat == list(range(lo, hi))
            seen += flat
        assert seen == list(range(wl["frames"]))
    one, _ = bench._batches_for_rank(bench.WORKLOADS["waymo_b4"], 0, 1)
    assert one == [[0, 1, 2, 3]]
