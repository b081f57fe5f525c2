"""Multi-GPU plumbing of the hot path (one process per GPU).

Reads are independent in both passes, so ranks take contiguous shards of the read-group index space and never
exchange reads; the only collective is the sum of the whitelist counts between the passes
(``pcl_counts_allreduce``, NCCL).  This module holds the host-side pieces: shard bounds and the exchange of
the 128-byte NCCL unique id over whatever ``torch.distributed`` backend the launcher initialised.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple


def shard_bounds(n_total: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous shards [g*N/G, (g+1)*N/G): concatenating the ranks' outputs restores the input order
    (the reference's rayon collect is order preserving, precellar/src/align/fastq.rs:341-348)."""
    return [(n_total * g // world, n_total * (g + 1) // world) for g in range(world)]


def exchange_unique_id(make_uid: Callable[[], bytes], rank: int, world: int, device: Optional[str] = None) -> Optional[bytes]:
    """Rank 0 creates the NCCL unique id, everybody receives it (torch.distributed broadcast)."""
    if world == 1:
        return None
    import torch
    import torch.distributed as dist
    dev = device or ("cuda" if dist.get_backend() == "nccl" else "cpu")
    t = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        t = torch.frombuffer(bytearray(make_uid()), dtype=torch.uint8).to(dev)
    dist.broadcast(t, 0)
    return bytes(t.cpu().numpy().tobytes())
