"""CPU, world_size 2 over gloo: the N > 1 host logic (leaf sharding, global playout ids, the single
counter all-reduce, winner gather).  The per-rank playouts run through the host build of the
product's ply code (tests/host_build), so the invariance "W ranks == 1 rank, bit for bit" is checked
end to end without a GPU."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, reps, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["OMP_NUM_THREADS"] = "2"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from conftest import HostBuild
    from gpu_ai_b200 import sharding
    hb = HostBuild()
    leaves = hb.gen_leaves(n_total, key=2016)          # every rank can regenerate any leaf: ids are global
    lo, hi = sharding.strong_shard(n_total, rank, world)
    # rep r of leaf i has pid = pid_base + r*n_total + i  -> run rep by rep with the shard's offset
    winners, counters = [], np.zeros(4, dtype=np.int64)
    for r in range(reps):
        w, p, _, c = hb.playouts(leaves[lo:hi], key=777, pid_base=sharding.shard_pid_base(1000 + r * n_total, lo), order=1)
        winners.append(w)
        counters += c.astype(np.int64)
    t = torch.from_numpy(counters.copy())
    sharding.allreduce_counters(t)
    g = sharding.gather_winners(torch.from_numpy(winners[0].copy()), n_total, rank, world)
    tmax = sharding.max_over_ranks(10.0 + rank)
    if rank == 0:
        np.savez(os.path.join(out_dir, "res.npz"), counters=t.numpy(), winners=g.numpy(), tmax=tmax)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_sharding_equals_single_rank(tmp_path, hostbuild):
    n_total, reps, world = 3001, 2, 2    # odd size: ragged shards
    mp.spawn(_worker, args=(world, _free_port(), n_total, reps, str(tmp_path)), nprocs=world, join=True)
    res = np.load(tmp_path / "res.npz")
    leaves = hostbuild.gen_leaves(n_total, key=2016)
    w, p, _, c = hostbuild.playouts(leaves, reps=reps, key=777, pid_base=1000, order=1)
    assert np.array_equal(res["counters"], c.astype(np.int64))
    assert np.array_equal(res["winners"], w[:n_total])
    assert float(res["tmax"]) == 11.0


def test_shard_arithmetic():
    from gpu_ai_b200 import sharding
    for n in (0, 1, 7, 1 << 20):
        for world in (1, 2, 4, 8):
            edges = [sharding.strong_shard(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
    assert sharding.weak_shard(1 << 20, 3) == (3 << 20, 4 << 20)
