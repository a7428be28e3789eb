"""N>1 host logic on CPU: two gloo ranks each search their contiguous query shard (with the oracle standing in for
the device, this box has no GPU) and rank 0 must end up with exactly the unsharded result."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, case, out_path):
    sys.path.insert(0, ROOT)
    from avxwindowfmindex_b200 import read_awfmi, sharding
    from oracle import harness
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    arrays = read_awfmi(os.path.join(GOLDEN, case + ".awfmi"))
    g = np.load(os.path.join(GOLDEN, case + ".npz"))
    offsets = g["offsets"]
    n = len(offsets) - 1
    a, b = sharding.shard_of(n, rank, world)
    local_offsets = offsets[a:b + 1] - offsets[a]
    local_letters = g["letters"][int(offsets[a]):int(offsets[b])]
    oracle = harness.Oracle(arrays)
    counts, _, _ = oracle.count(local_letters, local_offsets)
    hit, pos, _ = oracle.locate(local_letters, local_offsets)
    full_counts = sharding.gather_counts(torch.from_numpy(counts.astype(np.int32)), n)
    full_hit, full_pos = sharding.gather_hits(torch.from_numpy(hit.astype(np.int64)), torch.from_numpy(pos.astype(np.int64)), n)
    if rank == 0:
        np.savez(out_path, counts=full_counts.numpy(), hit=full_hit.numpy(), pos=full_pos.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_search_equals_unsharded(tmp_path):
    for case in ("nuc_k4_r4", "amino_k2_r3"):
        out = str(tmp_path / (case + ".npz"))
        mp.spawn(_worker, args=(2, _free_port(), case, out), nprocs=2, join=True)
        got = np.load(out)
        g = np.load(os.path.join(GOLDEN, case + ".npz"))
        assert np.array_equal(got["counts"].astype(np.uint32), g["counts"])
        assert np.array_equal(got["hit"].astype(np.uint64), g["hit_offsets"])
        assert np.array_equal(got["pos"].astype(np.uint64), g["positions"])


def test_shard_bounds_cover_everything():
    from avxwindowfmindex_b200 import sharding
    for n in (0, 1, 7, 100, 100_000_001):
        for world in (1, 2, 3, 8):
            b = sharding.shard_bounds(n, world)
            assert b[0] == 0 and b[-1] == n and all(x <= y for x, y in zip(b, b[1:]))
            assert max(y - x for x, y in zip(b, b[1:])) - min(y - x for x, y in zip(b, b[1:])) <= 1
