"""RHS-sharded multi-GPU driver: one process per GPU (torch.distributed, NCCL over NVLink), right-hand-side vectors
partitioned by rank, the triangular matrix broadcast once from its owner.  The right-hand sides are independent
(columns of B for side 'L', rows for side 'R'), so after the broadcast there is no further exchange: the reference has
no multi-device path at all, this is the one distributed strategy the B200 build adds.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple


def shard_range(m: int, world: int, rank: int, gran: int = 128) -> Tuple[int, int]:
    """Contiguous block of RHS vectors owned by `rank`: (first, count).  Blocks are multiples of `gran` (the GEMM N tile)
    except the last non-empty one; ranks beyond the data get (m, 0)."""
    if m < 0 or world < 1 or not (0 <= rank < world):
        raise ValueError("bad shard arguments")
    per = -(-m // world)
    per = -(-per // gran) * gran
    v0 = min(m, rank * per)
    return v0, max(0, min(per, m - v0))


def broadcast_matrix(A_storage, src: int = 0, group=None, chunks: int = 1):
    """Broadcast the contiguous storage of A from `src` to every rank (optionally in `chunks` pieces so a consumer can
    start on the leading block columns while the rest is still in flight).  Returns the list of async work handles."""
    import torch.distributed as dist

    flat = A_storage.view(-1)
    works = []
    step = -(-flat.numel() // max(1, chunks))
    for c in range(0, flat.numel(), step):
        works.append(dist.broadcast(flat[c:c + step], src=src, group=group, async_op=True))
    return works


def unified_rectrxm_sharded(side: str, uplo: str, transpose: str, alpha: float, func: str, A, B_local, src: int = 0, group=None,
                            solver: Optional[Callable] = None, need_broadcast: bool = True):
    """Every rank calls this with its own shard `B_local` (n x m_local for side 'L', m_local x n for side 'R') and a
    buffer `A` of the full order (contents significant on `src` only when `need_broadcast`).  In place on B_local."""
    import torch.distributed as dist

    if need_broadcast and dist.is_initialized() and dist.get_world_size(group) > 1:
        storage = A.t() if A.stride(0) == 1 and A.dim() == 2 and not A.is_contiguous() else A
        for w in broadcast_matrix(storage, src=src, group=group):
            w.wait()
    if solver is None:
        from . import unified_rectrxm as solver  # the CUDA library; there is no CPU fallback
    empty = B_local.shape[1] == 0 if side == "L" else B_local.shape[0] == 0
    if not empty:
        solver(side, uplo, transpose, alpha, func, A, B_local)
    return B_local
