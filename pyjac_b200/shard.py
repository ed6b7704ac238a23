"""State-batch sharding over the GPUs of one box.

Every state is independent, so the batch is cut into contiguous slices, one per rank; the
mechanism tables are replicated and the evaluation itself needs no collective.  The only
exchange is the optional gather of results to one rank (``gather_rows``), which works with
any ``torch.distributed`` backend (NCCL on the GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import List, Optional, Tuple


def partition(n: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous [start, stop) per rank: ceil(n / world) states each, the tail ranks may get
    fewer (or none)."""
    if n < 0 or world < 1:
        raise ValueError('bad partition request')
    per = -(-n // world) if n else 0
    return [(min(r * per, n), min((r + 1) * per, n)) for r in range(world)]


def my_slice(n: int, rank: int, world: int) -> Tuple[int, int]:
    return partition(n, world)[rank]


def gather_rows(local, n: int, dst: int = 0, group=None):
    """Gather the per-rank slices ``local[rows_r, width]`` of a row-per-state array of ``n``
    states to rank ``dst`` (returns the full array there, ``None`` elsewhere).  Slices are
    padded to the common slice length for the collective and trimmed afterwards."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    parts = partition(n, world)
    per = max(b - a for a, b in parts) if parts else 0
    a, b = parts[rank]
    assert local.shape[0] == b - a
    buf = local
    if b - a < per:
        buf = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        buf[:b - a] = local
    buf = buf.contiguous()
    outs: Optional[list] = [torch.empty_like(buf) for _ in range(world)] if rank == dst else None
    dist.gather(buf, outs, dst=dst, group=group)
    if rank != dst:
        return None
    return torch.cat([o[:pb - pa] for o, (pa, pb) in zip(outs, parts)], dim=0)
