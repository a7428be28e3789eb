"""Query sharding across ranks: the index is replicated, queries are split into contiguous shards,
and the only exchange is the gather of per-rank results onto one rank.  Backend-agnostic over torch.distributed:
NCCL with device tensors on the B200 box, gloo with CPU tensors in the CPU tests."""
import torch
import torch.distributed as dist


def shard_bounds(num_queries, world):
    """world+1 cut points; shard r = [bounds[r], bounds[r+1])  (query i -> rank floor(i*world/n), contiguous)"""
    return [num_queries * r // world for r in range(world + 1)]


def shard_of(num_queries, rank, world):
    b = shard_bounds(num_queries, world)
    return b[rank], b[rank + 1]


def gather_counts(local_counts, num_queries, dst=0, group=None):
    """Per-rank uint32/int32 count shards -> the full count array on rank `dst` (None elsewhere)."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    bounds = shard_bounds(num_queries, world)
    width = max(bounds[r + 1] - bounds[r] for r in range(world))
    padded = torch.zeros(width, dtype=local_counts.dtype, device=local_counts.device)
    padded[: local_counts.numel()] = local_counts
    bufs = [torch.empty_like(padded) for _ in range(world)] if rank == dst else None
    dist.gather(padded, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    return torch.cat([bufs[r][: bounds[r + 1] - bounds[r]] for r in range(world)])


def gather_hits(local_hit_offsets, local_positions, num_queries, dst=0, group=None):
    """Per-rank CSR (hit_offsets[shard+1], positions[hits] or payload[hits, k]) -> the global CSR on rank `dst`.
    Hit totals are exchanged first (all_gather of one int64), then the variable-length segments are gathered padded.
    A 2-D payload (e.g. position, contig, offset per hit) travels in ONE gather."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = local_positions.device
    hits = local_positions.shape[0]
    total = torch.tensor([hits], dtype=torch.int64, device=dev)
    totals = [torch.zeros_like(total) for _ in range(world)]
    dist.all_gather(totals, total, group=group)
    totals = [int(t) for t in torch.cat(totals).tolist()]
    counts = (local_hit_offsets[1:] - local_hit_offsets[:-1]).to(torch.int64)
    all_counts = gather_counts(counts, num_queries, dst, group)
    width = max(max(totals), 1)
    padded = torch.zeros((width,) + tuple(local_positions.shape[1:]), dtype=local_positions.dtype, device=dev)
    padded[:hits] = local_positions
    bufs = [torch.empty_like(padded) for _ in range(world)] if rank == dst else None
    dist.gather(padded, bufs, dst=dst, group=group)
    if rank != dst:
        return None, None
    positions = torch.cat([bufs[r][: totals[r]] for r in range(world)])
    hit_offsets = torch.zeros(num_queries + 1, dtype=torch.int64, device=dev)
    torch.cumsum(all_counts, 0, out=hit_offsets[1:])
    return hit_offsets, positions
