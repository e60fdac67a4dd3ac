"""World-size-2 gloo tests (CPU) of the multi-rank host logic: sharding by global board id, the all-reduce
plumbing used for gradients / baseline sums, and the oracle-level invariance 'two half-shards == one full
run' that the Philox keying provides."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import b2048
    from b2048 import dist as bd
    import oracle
    info = bd.init_distributed("gloo")
    assert (info.rank, info.world_size) == (rank, world)
    total = 1001
    lo, hi = bd.shard_range(total, rank, world)
    # 1. shards tile [0, total) without overlap
    owned = torch.zeros(total, dtype=torch.int64)
    owned[lo:hi] = 1
    bd.allreduce_sum_(owned)
    assert bool((owned == 1).all())
    # 2. gradient all-reduce: flat vector sums
    g = torch.full((71172,), float(rank + 1))
    bd.allreduce_sum_(g)
    assert float(g[0]) == sum(range(1, world + 1))
    # 3. baseline sums are plain sums: local (sum w, sum wv, sum wv^2, n) all-reduced == global statistics
    rng = np.random.default_rng(0)
    vals = rng.normal(size=total)
    mine = vals[lo:hi]
    st = torch.tensor([len(mine), mine.sum(), (mine ** 2).sum(), len(mine)], dtype=torch.float64)
    bd.allreduce_sum_(st)
    mean = st[1] / st[0]
    var = st[2] / st[0] - mean ** 2
    assert abs(float(mean) - vals.mean()) < 1e-12 and abs(float(var) - vals.var()) < 1e-12
    # 4. sharding invariance of the environment streams (CPU oracle, same definition as the kernels)
    seed, T = 99, 25
    cfg = oracle.make_cfg(action_mode="random_legal", auto_reset=True, max_steps=20, reward_mode="log2")
    st_local = oracle.reset_many(hi - lo, seed, lo, 0)
    for t in range(1, T + 1):
        oracle.step_many(st_local, cfg, seed, lo, t)
    boards = torch.zeros(total, dtype=torch.int64)
    boards[lo:hi] = torch.from_numpy(st_local["board"].view(np.int64))
    bd.allreduce_sum_(boards)
    if rank == 0:
        full = oracle.reset_many(total, seed, 0, 0)
        for t in range(1, T + 1):
            oracle.step_many(full, cfg, seed, 0, t)
        assert (boards.numpy().view(np.uint64) == full["board"]).all()
    # 4b. episode rank weights: all-gathered global ranks == the single-process weights of the same episodes
    rew = torch.from_numpy(np.random.default_rng(5).normal(size=total) * 50)
    w_mine = bd.episode_rank_weights(rew[lo:hi].clone(), [0.0, 0.5, 1.0, 2.5], info)
    w_full = bd.episode_rank_weights(rew, [0.0, 0.5, 1.0, 2.5])
    assert torch.allclose(w_mine, w_full[lo:hi])
    # 5. max-over-ranks timing reduction used by bench.py
    tm = torch.tensor([float(rank + 1)], dtype=torch.float64)
    bd.allreduce_max_(tm)
    assert float(tm) == world
    dist.barrier()
    dist.destroy_process_group()
    out.put((rank, "ok"))


def test_two_rank_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
        assert p.exitcode == 0
    got = sorted(q.get(timeout=5) for _ in range(2))
    assert got == [(0, "ok"), (1, "ok")]


def test_shard_range_edges():
    sys.path.insert(0, ROOT)
    from b2048 import dist as bd
    for total in (0, 1, 7, 64, 1 << 26):
        for world in (1, 2, 3, 8):
            r = [bd.shard_range(total, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == total
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1
