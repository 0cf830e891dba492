"""Multi-GPU plumbing of the CCD step: one process per GPU, torch.distributed (NCCL on the GPUs, gloo in CPU tests).

The path shards without a data-path collective: positions and faces are replicated (every rank uploads its slice and
all-gathers, or receives the same q0/q1 — `gather_positions` / `broadcast_positions`), every rank builds the same LBVH,
and rank r intersects it, emits and tests only for the stencils it owns: the vertices (VF) and unique edges (EE) anchored
in a contiguous range of the faces' Morton order — a range of LBVH subtrees, i.e. a region of space.
The only exchange is the step summary, ONE small all-gather per step (`exchange_step`): earliest TOI, hit / stencil
counts, and each rank's load profile (stencils per bucket of sorted positions), from which every rank derives the same
ownership ranges for the next step (`balanced_bounds`) — the rebalancing SURVEY.md section 8(e) asks for.
"""
import numpy as np
import torch
import torch.distributed as dist


# fallback cost model when no stage times are available: cost of owning one sorted position (box tests and tree
# intersection around that face) in units of the cost of one stencil (emission + narrowphase)
VERTEX_WEIGHT = 2.0


def shard_range(n, rank, world):
    """Contiguous range [begin, end) of rank `rank` over n equally weighted items."""
    return (n * rank) // world, (n * (rank + 1)) // world


def broadcast_positions(q0, q1, src=0, group=None):
    """Replicated-position broadcast (48 B per vertex for a single step).  In place on tensors of any backend."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(q0, src, group=group)
        dist.broadcast(q1, src, group=group)
    return q0, q1


def reduce_step_summary(earliest_toi, n_hits, n_stencils, device="cpu", group=None):
    """Global (earliest TOI, hits, stencils) from the per-rank values with a single MAX all-reduce.

    Slot layout: [ -toi | hits of rank 0..W-1 | stencils of rank 0..W-1 ]; counts are non-negative, so MAX over ranks
    of a vector that is zero everywhere except the rank's own slots gathers them, and the sums are taken locally.
    Counts up to 2^53 are exact in float64.
    """
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return float(earliest_toi), int(n_hits), int(n_stencils)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    buf = torch.zeros(1 + 2 * world, dtype=torch.float64, device=device)
    buf[0] = -earliest_toi if np.isfinite(earliest_toi) else -np.inf
    buf[1 + rank] = float(n_hits)
    buf[1 + world + rank] = float(n_stencils)
    dist.all_reduce(buf, op=dist.ReduceOp.MAX, group=group)
    out = buf.cpu().numpy()
    toi = -out[0]
    return (float(toi) if np.isfinite(toi) else float("inf")), int(out[1:1 + world].sum()), int(out[1 + world:].sum())


_FIRST_CACHE = {}


def _bucket_first(n_items, nb):
    """first item of each bucket b = 0..nb (item i falls in bucket i * nb // n_items)"""
    key = (int(n_items), int(nb))
    f = _FIRST_CACHE.get(key)
    if f is None:
        f = -(-(np.arange(nb + 1, dtype=np.int64) * int(n_items)) // int(nb))
        _FIRST_CACHE.clear()
        _FIRST_CACHE[key] = f
    return f


def balanced_bounds(hist, n_items, world):
    """world+1 ascending item bounds such that every rank gets about the same load, from a load profile.

    `hist[b]` is the load of the items i with i * len(hist) // n_items == b; inside a bucket the load is taken as
    uniform.  Pure function of its arguments: every rank that calls it with the same reduced histogram gets the same
    partition."""
    hist = np.asarray(hist, dtype=np.float64)
    nb = len(hist)
    total = float(hist.sum())
    bounds = np.zeros(world + 1, dtype=np.int64)
    bounds[world] = n_items
    if total <= 0 or n_items <= 0:
        bounds[1:world] = (int(n_items) * np.arange(1, world, dtype=np.int64)) // world
        return bounds.astype(np.int32)
    cum = np.concatenate([[0.0], np.cumsum(hist)])
    first = _bucket_first(n_items, nb)
    targets = total * np.arange(1, world, dtype=np.float64) / world
    b = np.clip(np.searchsorted(cum, targets, side="right") - 1, 0, nb - 1)
    hb = hist[b]
    frac = np.where(hb > 0, (targets - cum[b]) / np.where(hb > 0, hb, 1.0), 0.0)
    raw = np.rint(first[b] + frac * (first[b + 1] - first[b])).astype(np.int64)
    bounds[1:world] = np.minimum(np.maximum.accumulate(raw), n_items)
    return bounds.astype(np.int32)


def gather_positions(d_q0, d_q1, h_q0, h_q1, rank, world, group=None):
    """End-to-end input path of a sharded job: every rank copies only ITS 1/world slice of the step's positions from (pinned)
    host memory and the slices are all-gathered over NVLink (NCCL) into the replicated device arrays — the host-to-device
    traffic of the job is one copy of the positions, not `world` copies.  d_q* / h_q*: flat float64 tensors of equal
    length (device / pinned host)."""
    n = d_q0.numel()
    if not (dist.is_available() and dist.is_initialized()) or world == 1:
        d_q0.copy_(h_q0, non_blocking=True)
        d_q1.copy_(h_q1, non_blocking=True)
        return
    per = -(-n // world)
    for d, h in ((d_q0, h_q0), (d_q1, h_q1)):
        if n % world == 0:
            a, b = rank * per, (rank + 1) * per
            d[a:b].copy_(h[a:b], non_blocking=True)
            dist.all_gather_into_tensor(d, d[a:b], group=group)
        else:      # ragged tail: gather padded slices, then trim
            buf = torch.zeros(per * world, dtype=d.dtype, device=d.device)
            a, b = rank * per, min((rank + 1) * per, n)
            if b > a:
                buf[a:b].copy_(h[a:b], non_blocking=True)
            dist.all_gather_into_tensor(buf, buf[rank * per:(rank + 1) * per].clone(), group=group)
            d.copy_(buf[:n])


def _exchange(world, rank, head_vals, vf_hist, ee_hist, npos, device, group, rebalance):
    """The all-gather and the arithmetic of exchange_step, free of the context: returns (toi, hits, stencils, bounds or None)."""
    nb = len(vf_hist)
    nh = len(head_vals)
    mine = np.empty(nh + 2 * nb, dtype=np.float64)      # counts up to 2^53 are exact in float64
    mine[:nh] = head_vals
    mine[nh:nh + nb] = vf_hist
    mine[nh + nb:] = ee_hist
    send = torch.from_numpy(mine).to(device)
    recv = torch.empty(world * len(mine), dtype=torch.float64, device=device)
    dist.all_gather_into_tensor(recv, send, group=group)
    allr = recv.cpu().numpy().reshape(world, -1)
    pb = None
    if rebalance:
        tot_vf, tot_ee, tot_owned = allr[:, 2].sum(), allr[:, 3].sum(), allr[:, 4].sum()
        emit_per_stencil = allr[:, 6].sum() / max(tot_vf + tot_ee, 1.0)
        a = allr[:, 5].sum() / max(tot_owned, 1.0)
        b = allr[:, 7].sum() / max(tot_vf, 1.0) + emit_per_stencil
        c = allr[:, 8].sum() / max(tot_ee, 1.0) + emit_per_stencil
        if not (a > 0 and b > 0 and c > 0):      # no timings (CPU tests): the fixed model
            a, b, c = VERTEX_WEIGHT, 1.0, 1.0
        items = np.diff(_bucket_first(npos, nb)).astype(np.float64)
        load = b * allr[:, nh:nh + nb].sum(axis=0) + c * allr[:, nh + nb:].sum(axis=0) + a * items
        pb = balanced_bounds(load, npos, world)
    toi = allr[:, 0].min()
    return (float(toi) if np.isfinite(toi) else float("inf")), int(allr[:, 1].sum()), int(allr[:, 2].sum() + allr[:, 3].sum()), pb


def _exchange_head(ctx, world, rank, earliest_toi, n_hits, n_vf, n_ee, npos, stage_ms):
    t_item = stage_ms.get("traverse_exact", 0.0) + stage_ms.get("adjacency", 0.0)
    t_emit = stage_ms.get("emit_count", 0.0) + stage_ms.get("emit_write", 0.0)
    part = getattr(ctx, "shard_partition", None)
    owned = float(part[rank + 1] - part[rank]) if part is not None and len(part) == world + 1 else float(npos) / world
    return [earliest_toi if np.isfinite(earliest_toi) else np.inf, float(n_hits), float(n_vf), float(n_ee), owned, t_item, t_emit,
            stage_ms.get("np_vf", 0.0), stage_ms.get("np_ee", 0.0)]


def exchange_step(ctx, earliest_toi, n_hits, n_vf, n_ee=0, device="cpu", group=None, rebalance=True, stage_ms=None):
    """The step's only exchange: ONE all-gather of a small vector per rank — earliest TOI, hit / stencil counts, the rank's
    measured stage times and its load profile (stencils per bucket of sorted positions).

    Returns the global (earliest TOI, hits, stencils).  With `rebalance`, every rank also installs the ownership ranges
    for its next step (ccd_set_shard_partition), balanced on the summed profile under a cost model whose three
    coefficients are re-measured every step from the gathered times: cost of a rank = a * positions owned (box tests and
    tree intersection around them) + b * VF stencils + c * EE stencils (emission + narrowphase).
    `stage_ms`: this rank's ccd_stage_times() of the step (taken from ctx when omitted)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return float(earliest_toi), int(n_hits), int(n_vf + n_ee)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    vf_hist, ee_hist, npos = ctx.shard_histogram()
    if stage_ms is None:
        stage_ms = ctx.stage_times()
    head = _exchange_head(ctx, world, rank, earliest_toi, n_hits, n_vf, n_ee, npos, stage_ms)
    toi, hits, stencils, pb = _exchange(world, rank, head, vf_hist, ee_hist, npos, device, group, rebalance)
    if pb is not None:
        ctx.set_shard_partition(pb)
        ctx.shard_partition = list(map(int, pb))
    return toi, hits, stencils


class StepExchange(object):
    """exchange_step, pipelined: the all-gather of step k runs in a worker thread BESIDE step k + 1 (the step call blocks the
    calling thread inside the C library with the GIL released), and the ranges it yields are installed before step k + 2 —
    the load profile is one step stale, the GPUs never wait for the exchange or for each other between steps.  The global
    summary of step k is available from result() once the following submit() — or result() itself — has joined the worker.
    Every rank issues its collectives from its worker in step order, so they match across ranks."""

    def __init__(self, ctx, device="cpu", group=None, rebalance=True, cuda_index=None):
        import threading
        self._threading = threading
        self.ctx, self.device, self.group, self.rebalance, self.cuda_index = ctx, device, group, rebalance, cuda_index
        self.active = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        self.thread, self.out, self.last = None, None, None

    def _join(self):
        if self.thread is not None:
            self.thread.join()
            self.thread = None
            toi, hits, stencils, pb = self.out
            if pb is not None:
                self.ctx.set_shard_partition(pb)
                self.ctx.shard_partition = list(map(int, pb))
            self.last = (toi, hits, stencils)

    def submit(self, earliest_toi, n_hits, n_vf, n_ee, stage_ms):
        """Call right after a step: takes the step's load profile off the context, joins the exchange of the step before
        (installing its ranges for the NEXT step) and starts this step's exchange in the background."""
        if not self.active:
            self.last = (float(earliest_toi), int(n_hits), int(n_vf + n_ee))
            return
        world, rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        vf_hist, ee_hist, npos = self.ctx.shard_histogram()
        vf_hist, ee_hist = np.array(vf_hist), np.array(ee_hist)      # the context reuses its buffers in the next step
        head = _exchange_head(self.ctx, world, rank, earliest_toi, n_hits, n_vf, n_ee, npos, stage_ms)
        self._join()

        def run():
            if self.cuda_index is not None:
                torch.cuda.set_device(self.cuda_index)      # the current device is per thread
            self.out = _exchange(world, rank, head, vf_hist, ee_hist, npos, self.device, self.group, self.rebalance)
        self.thread = self._threading.Thread(target=run)
        self.thread.start()

    def result(self):
        """Global (earliest TOI, hits, stencils) of the last submitted step."""
        self._join()
        return self.last


