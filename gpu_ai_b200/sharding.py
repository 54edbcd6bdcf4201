"""Leaf sharding across ranks (one process per GPU) -- the only "parallelism" this path needs.

Playouts are independent, so there is no data-path collective: rank r of W plays a contiguous leaf
range and the four win counters {draws, P1, P2, plies} are combined by ONE 32-byte all-reduce per
iteration (NCCL on GPUs; the same code runs over gloo in the CPU tests).  Playout ids are GLOBAL
(pid = pid_base + rep * n_total + leaf), so winners are bit-identical for W = 1, 2, 4, 8.
Mirrors the in-process sharding of b2p_run_packed (gpu_ai_b200/csrc/api.cu: shard_of).
"""
import torch
import torch.distributed as dist


def strong_shard(n_total, rank, world):
    """Contiguous split of a fixed batch: rank r owns [lo, hi)."""
    return n_total * rank // world, n_total * (rank + 1) // world


def weak_shard(leaves_per_rank, rank):
    """Fixed work per rank (bench.py): rank r owns leaves [r*n, (r+1)*n) of the D_ref stream."""
    return leaves_per_rank * rank, leaves_per_rank * (rank + 1)


def shard_pid_base(pid_base, lo):
    """pid_base to hand to the engine for a shard starting at global leaf `lo` (rep_stride = n_total)."""
    return pid_base + lo


def allreduce_counters(counters):
    """Sum the 4 x int64 counters over all ranks in place (no-op without a process group)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(counters, op=dist.ReduceOp.SUM)
    return counters


def gather_winners(local_winners, n_total, rank, world):
    """All-gather per-leaf winners of a strong-sharded batch into global leaf order (reps == 1)."""
    if not (dist.is_available() and dist.is_initialized()) or world == 1:
        return local_winners
    sizes = [strong_shard(n_total, r, world) for r in range(world)]
    width = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros(width, dtype=local_winners.dtype, device=local_winners.device)
    pad[: local_winners.numel()] = local_winners
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    return torch.cat([p[: hi - lo] for p, (lo, hi) in zip(parts, sizes)])


def max_over_ranks(value, device="cpu"):
    """Timing rule: a multi-GPU number is the max over ranks."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
