"""Host-side logic of the multi-GPU mode: who owns which transfers, and how per-rank results combine.

The path shards by independent buffer ranges: rank r of N takes the contiguous
transfers perseus_gpu_shard_range() assigns it and nothing is exchanged on the data path.  The only
collectives are on the control path — a barrier around the timed region, MAX over ranks of the
device-measured time, and a modular SUM of the per-shard checksums — so they run on whatever
backend the process group has (NCCL on the GPU box, gloo in the CPU tests).
"""
from __future__ import annotations

MASK64 = (1 << 64) - 1


def rank_shard(pg, total_buffers: int, world: int, rank: int) -> tuple[int, int]:
    """(first transfer, number of transfers) owned by `rank`; thin wrapper over the C ABI."""
    return pg.shard_range(total_buffers, world, rank)


def _device_for(dist):
    import torch
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


def allreduce_max(x: float) -> float:
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(x)
    t = torch.tensor([x], dtype=torch.float64, device=_device_for(dist))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def allreduce_sum_u64(x: int) -> int:
    """Sum modulo 2^64 of one unsigned 64-bit value per rank (the shard checksums).  Sent as four
    16-bit limbs in int64 lanes so no backend ever sees an overflow."""
    import torch
    import torch.distributed as dist
    x &= MASK64
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return x
    limbs = torch.tensor([(x >> (16 * k)) & 0xFFFF for k in range(4)], dtype=torch.int64, device=_device_for(dist))
    dist.all_reduce(limbs, op=dist.ReduceOp.SUM)
    return sum(int(v) << (16 * k) for k, v in enumerate(limbs.tolist())) & MASK64


def allgather_float(x: float) -> list[float]:
    """Every rank's value, in rank order (e.g. the host-link rates that weight the shards of a host-fed recording)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [float(x)]
    mine = torch.tensor([x], dtype=torch.float64, device=_device_for(dist))
    out = [torch.zeros_like(mine) for _ in range(dist.get_world_size())]
    dist.all_gather(out, mine)
    return [float(t.item()) for t in out]


def observed_rates(units: float, seconds: float) -> list[float]:
    """Every rank's units / seconds over the pass it just finished, in rank order: the shard weights of the next pass
    (perseus_gpu_shard_range_weighted).  A rank that finishes early leaves the shared host path to the others, which flatters
    THEIR rates, so weights taken from an unbalanced pass over-feed the slow ranks; once the shards finish together the
    observed rates are the steady-state ones.  Two or three passes settle it."""
    return allgather_float(units / max(seconds, 1e-9))


def gather_ranges(first: int, count: int) -> list[tuple[int, int]]:
    """Every rank's (first, count), in rank order."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [(first, count)]
    mine = torch.tensor([first, count], dtype=torch.int64, device=_device_for(dist))
    out = [torch.zeros_like(mine) for _ in range(dist.get_world_size())]
    dist.all_gather(out, mine)
    return [(int(t[0]), int(t[1])) for t in out]
