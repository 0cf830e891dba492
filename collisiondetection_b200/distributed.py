"""Multi-GPU plumbing of the CCD step: one process per GPU, torch.distributed (NCCL on the GPUs, gloo in CPU tests).

The path shards without a data-path collective: positions and faces are replicated (every rank uploads or receives
the same q0/q1 — `broadcast_positions`), every rank builds the same LBVH, and rank r emits and tests only the stencils
it owns: a contiguous vertex range for VF and unique-edge range for EE, balanced by stencil count inside
ccd_step_device (`shard_range` is the plain index split used for other per-item work).
The only exchange is the step summary: earliest TOI (min) and hit / stencil counts (sum), fused into ONE all-reduce
by carrying the minimum as a negated maximum next to sums that are kept on their own rank's slot.
"""
import numpy as np
import torch
import torch.distributed as dist


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
