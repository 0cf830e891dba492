"""Multi-GPU plumbing of the CCD step: one process per GPU, torch.distributed (NCCL on the GPUs, gloo in CPU tests).

The path shards without a data-path collective: positions and faces are replicated (every rank uploads or receives
the same q0/q1 — `broadcast_positions`), every rank builds the same LBVH, and rank r traverses it, emits and tests only
for the stencils it owns: a contiguous vertex range (VF) and unique-edge range (EE).
The only exchange is the step summary, ONE small all-gather per step (`exchange_step`): earliest TOI, hit / stencil
counts, and each rank's load profile (stencils per bucket of vertex / edge ids), from which every rank derives the same
ownership ranges for the next step (`balanced_bounds`) — the rebalancing SURVEY.md section 8(e) asks for.
"""
import numpy as np
import torch
import torch.distributed as dist


# cost of owning a vertex (box tests and LBVH traversal for the ~2 faces around it, both directions) in units of the
# cost of one stencil (emission + narrowphase); measured on B200 at the 4M-triangle cloth: ~3.5 ns vs ~0.9 ns
VERTEX_WEIGHT = 4.0


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


def balanced_bounds(hist, n_items, world):
    """world+1 ascending item bounds such that every rank gets about the same load, from a load profile.

    `hist[b]` is the load of the items i with i * len(hist) // n_items == b; inside a bucket the load is taken as
    uniform.  Pure function of its arguments: every rank that calls it with the same reduced histogram gets the same
    partition."""
    hist = np.asarray(hist, dtype=np.float64)
    nb = len(hist)
    total = hist.sum()
    bounds = np.zeros(world + 1, dtype=np.int64)
    bounds[world] = n_items
    if total <= 0 or n_items <= 0:
        for r in range(1, world):
            bounds[r] = (n_items * r) // world
        return bounds.astype(np.int32)
    cum = np.concatenate([[0.0], np.cumsum(hist)])
    # first item of bucket b
    first = [-(-(b * n_items) // nb) for b in range(nb + 1)]
    for r in range(1, world):
        target = total * r / world
        b = int(np.searchsorted(cum, target, side="right") - 1)
        b = min(max(b, 0), nb - 1)
        frac = (target - cum[b]) / hist[b] if hist[b] > 0 else 0.0
        lo, hi = first[b], first[b + 1]
        bounds[r] = min(max(int(round(lo + frac * (hi - lo))), bounds[r - 1]), n_items)
    return bounds.astype(np.int32)


def exchange_step(ctx, earliest_toi, n_hits, n_stencils, device="cpu", group=None, rebalance=True):
    """The step's only exchange: ONE all-gather of [toi, hits, stencils, vf load profile, ee load profile] per rank.

    Returns the global (earliest TOI, hits, stencils).  With `rebalance`, every rank also installs the ownership ranges
    balanced on the summed load profile for its next step (ccd_set_shard_partition)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return float(earliest_toi), int(n_hits), int(n_stencils)
    world = dist.get_world_size(group)
    vf_hist, ee_hist, nv, ne = ctx.shard_histogram()
    nb = len(vf_hist)
    mine = np.empty(3 + 2 * nb, dtype=np.float64)      # counts up to 2^53 are exact in float64
    mine[0] = earliest_toi if np.isfinite(earliest_toi) else np.inf
    mine[1], mine[2] = float(n_hits), float(n_stencils)
    mine[3:3 + nb] = vf_hist
    mine[3 + nb:] = ee_hist
    send = torch.from_numpy(mine).to(device)
    recv = torch.empty(world * len(mine), dtype=torch.float64, device=device)
    dist.all_gather_into_tensor(recv, send, group=group)
    allr = recv.cpu().numpy().reshape(world, -1)
    if rebalance:
        # cost model of a vertex range: its stencils (emission + narrowphase) plus, per vertex, the traversal of the faces
        # around it; both profiles are over vertex ids, and the edge range follows the vertex range (shard_edge_bounds)
        items = np.diff([-(-(b * nv) // nb) for b in range(nb + 1)]).astype(np.float64)
        load = allr[:, 3:3 + nb].sum(axis=0) + allr[:, 3 + nb:].sum(axis=0) + VERTEX_WEIGHT * items
        vb = balanced_bounds(load, nv, world)
        ctx.set_shard_partition(vb, ctx.shard_edge_bounds(vb))
    toi = allr[:, 0].min()
    return (float(toi) if np.isfinite(toi) else float("inf")), int(allr[:, 1].sum()), int(allr[:, 2].sum())
