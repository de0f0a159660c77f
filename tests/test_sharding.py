"""The N > 1 path on CPU: chain sharding and the counter / histogram / timing reductions over a world_size-2 gloo
group (the same code runs over NCCL on the GPUs, see bench.py)."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from jellyfysh_b200 import sharding


def test_shards_are_disjoint_and_cover_all_chains():
    for world in (1, 2, 4, 8):
        shards = [sharding.chain_shard(rank, world, 4096) for rank in range(world)]
        covered = np.concatenate([np.arange(first, first + count) for first, count in shards])
        assert np.array_equal(covered, np.arange(world * 4096))
        split = sharding.split_chains(4096 + 3, world)
        assert sum(count for _, count in split) == 4099
        assert all(split[r][0] + split[r][1] == split[r + 1][0] for r in range(world - 1))
        assert max(c for _, c in split) - min(c for _, c in split) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        first, count = sharding.chain_shard(rank, world, 8)
        stats = {key: (rank + 1) * (k + 1) for k, key in enumerate(sharding.COUNTER_KEYS)}
        stats["events"] = count * 100 + first
        total = sharding.reduce_counters(stats)
        histogram = np.zeros(16, dtype=np.int64)
        histogram[first % 16] = 5
        histogram[3] += rank + 1
        reduced = sharding.reduce_histogram(histogram)
        slowest = sharding.reduce_max([1.0 + rank, 10.0 - rank])
        results[rank] = (total, reduced.tolist(), slowest.tolist())
    finally:
        dist.destroy_process_group()


def test_reductions_over_gloo_world_size_2():
    world = 2
    manager = mp.Manager()
    results = manager.dict()
    mp.spawn(_worker, args=(world, _free_port(), results), nprocs=world, join=True)
    assert len(results) == world
    for rank in range(world):
        total, histogram, slowest = results[rank]
        assert total["events"] == (8 * 100 + 0) + (8 * 100 + 8)
        assert total["pair_events"] == (1 + 2) * 2 and total["capacity_errors"] == (1 + 2) * 9
        expected = np.zeros(16, dtype=np.int64)
        expected[0] += 5
        expected[8] += 5
        expected[3] += 3
        assert histogram == expected.tolist()
        assert slowest == [2.0, 10.0]


def test_reductions_without_a_process_group_are_identities():
    stats = {key: k for k, key in enumerate(sharding.COUNTER_KEYS)}
    assert sharding.reduce_counters(stats) == stats
    assert sharding.reduce_max([1.5]).tolist() == [1.5]
