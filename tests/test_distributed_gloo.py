"""world_size-2 (and 3) gloo runs of the host-side multi-GPU logic on CPU: ownership ranges partition the stencil
lists exactly, and the fused summary all-reduce reproduces the single-process result."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, golden

sys.path.insert(0, ROOT)


def _worker(rank, world, port_no, vf, ee, vf_hit, vf_toi, ee_hit, ee_toi, V, edges, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from collisiondetection_b200 import distributed as D
    # the stencils this rank owns: VF by vertex range, EE by rank range of the first edge (ids are lexicographic)
    v0, v1 = D.shard_range(V, rank, world)
    e0, e1 = D.shard_range(len(edges), rank, world)
    own_vf = (vf[:, 0] >= v0) & (vf[:, 0] < v1)
    first_edge = np.searchsorted(edges, vf_key(ee[:, 0], ee[:, 1]))
    own_ee = (first_edge >= e0) & (first_edge < e1)
    hits = int(vf_hit[own_vf].sum() + ee_hit[own_ee].sum())
    tois = np.concatenate([vf_toi[own_vf][vf_hit[own_vf] > 0], ee_toi[own_ee][ee_hit[own_ee] > 0]])
    earliest = tois.min() if tois.size else np.inf
    q = torch.arange(6, dtype=torch.float64) * (1.0 if rank == 0 else 0.0)
    q2 = q.clone()
    D.broadcast_positions(q, q2)
    assert torch.equal(q, torch.arange(6, dtype=torch.float64))
    toi, nh, ns = D.reduce_step_summary(earliest, hits, int(own_vf.sum() + own_ee.sum()))
    results[rank] = (toi, nh, ns, int(own_vf.sum()), int(own_ee.sum()))
    dist.destroy_process_group()


def vf_key(a, b):
    return a.astype(np.int64) << 32 | b.astype(np.int64)


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_summary_matches_whole(world):
    g = golden("alec_prob11_835.npz")
    vf, ee = g["ref_vf"], g["ref_ee"]
    V = len(g["q0"])
    f = g["faces"]
    he = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]])
    edges = np.unique(vf_key(he.min(1), he.max(1)))
    mgr = mp.get_context("spawn").Manager()
    results = mgr.dict()
    port_no = 29600 + world
    mp.spawn(_worker, args=(world, port_no, vf, ee, g["ref_vf_hit"], g["ref_vf_toi"], g["ref_ee_hit"], g["ref_ee_toi"], V, edges, results),
             nprocs=world, join=True)
    hits = np.concatenate([g["ref_vf_toi"][g["ref_vf_hit"] > 0], g["ref_ee_toi"][g["ref_ee_hit"] > 0]])
    for r in range(world):
        toi, nh, ns, _, _ = results[r]
        assert toi == hits.min()
        assert nh == int(g["ref_vf_hit"].sum() + g["ref_ee_hit"].sum())
        assert ns == len(vf) + len(ee)
    assert sum(results[r][3] for r in range(world)) == len(vf)
    assert sum(results[r][4] for r in range(world)) == len(ee)


def test_shard_ranges_partition():
    from collisiondetection_b200.distributed import shard_range
    for n in (0, 1, 7, 1000, 2002225):
        for world in (1, 2, 3, 4, 8):
            ranges = [shard_range(n, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in ranges]
            assert max(sizes) - min(sizes) <= 1


def test_balanced_bounds():
    from collisiondetection_b200.distributed import balanced_bounds
    rng = np.random.default_rng(3)
    for n_items in (5, 1000, 2002225):
        for world in (1, 2, 3, 8):
            nb = 1024
            bucket = (np.arange(n_items, dtype=np.int64) * nb) // n_items
            load = rng.integers(0, 20, size=n_items) * (np.arange(n_items) > n_items // 3)      # one third of the ids idle
            hist = np.bincount(bucket, weights=load, minlength=nb)
            b = balanced_bounds(hist, n_items, world)
            assert b[0] == 0 and b[-1] == n_items and np.all(np.diff(b) >= 0) and len(b) == world + 1
            if n_items >= 1000:
                per = np.array([load[b[r]:b[r + 1]].sum() for r in range(world)], dtype=np.float64)
                assert per.max() <= 1.05 * per.mean() + 40, (n_items, world, per)
    # empty profile: equal index split
    assert list(balanced_bounds(np.zeros(16), 100, 4)) == [0, 25, 50, 75, 100]


class _FakeCtx(object):
    """Stands in for api.Context in the CPU test of the exchange: a fixed load profile per rank, records the partition."""
    SHARD_BUCKETS = 1024

    def __init__(self, rank, world, npos):
        self.rank, self.world, self.npos = rank, world, npos
        self.partition = None

    def shard_histogram(self):
        vf = np.zeros(self.SHARD_BUCKETS, np.int64)
        ee = np.zeros(self.SHARD_BUCKETS, np.int64)
        lo, hi = (self.SHARD_BUCKETS * self.rank) // self.world, (self.SHARD_BUCKETS * (self.rank + 1)) // self.world
        vf[lo:hi] = 10 * (self.rank + 1)      # later ranks are heavier
        ee[lo:hi] = 7
        return vf, ee, self.npos

    def set_shard_partition(self, pb):
        self.partition = list(map(int, pb))


def _exchange_worker(rank, world, port_no, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from collisiondetection_b200 import distributed as D
    ctx = _FakeCtx(rank, world, 100000)
    toi, nh, ns = D.exchange_step(ctx, 0.25 + 0.1 * rank if rank else float("inf"), 10 + rank, 600 * (rank + 1), 400 * (rank + 1), stage_ms={})
    # the end-to-end input path: every rank contributes its slice, all ranks end with the whole arrays
    n = 1000 + rank * 0 + (3 if world == 3 else 0)
    h0 = torch.arange(n, dtype=torch.float64)
    h1 = h0 * 2.0
    d0, d1 = torch.zeros(n, dtype=torch.float64), torch.zeros(n, dtype=torch.float64)
    D.gather_positions(d0, d1, h0, h1, rank, world)
    assert torch.equal(d0, h0) and torch.equal(d1, h1)
    results[rank] = (toi, nh, ns, ctx.partition)
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_exchange_step_gathers_and_rebalances(world):
    mgr = mp.get_context("spawn").Manager()
    results = mgr.dict()
    mp.spawn(_exchange_worker, args=(world, 29650 + world, results), nprocs=world, join=True)
    parts = {results[r][3] and tuple(results[r][3]) for r in range(world)}
    assert len(parts) == 1, "ranks derived different partitions"
    pb = results[0][3]
    assert pb[0] == 0 and pb[-1] == 100000
    # heavier late ranks -> their position ranges shrink
    assert pb[1] > 100000 // world
    for r in range(world):
        toi, nh, ns, _ = results[r]
        assert toi == (0.35 if world > 1 else float("inf"))
        assert nh == sum(10 + k for k in range(world)) and ns == sum(1000 * (k + 1) for k in range(world))


def _pipelined_worker(rank, world, port_no, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from collisiondetection_b200 import distributed as D
    ctx = _FakeCtx(rank, world, 100000)
    x = D.StepExchange(ctx)
    seen = []
    for step in range(3):
        x.submit(0.5 - 0.1 * step + 0.01 * rank, 10 * (step + 1) + rank, 600 * (rank + 1), 400 * (rank + 1), {})
        seen.append(ctx.partition)      # the ranges of step k are installed one submit later
    last = x.result()
    results[rank] = (last, seen, ctx.partition)
    dist.destroy_process_group()


def test_pipelined_exchange_is_one_step_stale_and_consistent():
    """StepExchange: the exchange of step k overlaps step k + 1; its ranges arrive before step k + 2, the same on every rank,
    and result() returns the summary of the last submitted step."""
    world = 2
    mgr = mp.get_context("spawn").Manager()
    results = mgr.dict()
    mp.spawn(_pipelined_worker, args=(world, 29671, results), nprocs=world, join=True)
    for r in range(world):
        (toi, nh, ns), seen, final = results[r]
        assert seen[0] is None                      # nothing installed by the first submit
        assert seen[1] is not None and seen[1] == seen[2] == final      # same profile every step -> same ranges
        assert abs(toi - 0.3) < 1e-12 and nh == 30 + 31 and ns == 1000 * 3
    assert results[0][2] == results[1][2]
