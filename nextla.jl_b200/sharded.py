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


def panel_geometry(n: int, panels: int, gran: int = 128) -> Tuple[int, int]:
    """(panel_cols, n_panels): about `panels` column panels of A, a multiple of `gran` columns wide."""
    if n <= 0:
        return gran, 0
    pc = -(-n // max(1, panels))
    pc = -(-pc // gran) * gran
    return pc, -(-n // pc)


def broadcast_panels(A_storage, order, panel_cols: int, src: int = 0, group=None, on_panel: Optional[Callable] = None):
    """Broadcast A panel by panel in the given order.  `A_storage` is the contiguous (n x n) storage of a column-major A, so
    column panel p of A is the row block [p*panel_cols, (p+1)*panel_cols) of the storage.  `on_panel(p)` runs after each
    panel has been enqueued (the CUDA path records an event there)."""
    import torch.distributed as dist

    n = A_storage.shape[0]
    for p in order:
        dist.broadcast(A_storage[p * panel_cols:min(n, (p + 1) * panel_cols)], src=src, group=group)
        if on_panel is not None:
            on_panel(p)


_side_streams = {}
_stage = {}


def broadcast_panel_triangle(A_store, uplo: str, p: int, pc: int, src: int, group, stream, handle):
    """Broadcast only the referenced part of column panel p of A (CUDA): rows [p*pc, n) of a lower panel, [0, (p+1)*pc) of an upper
    one -- a trapezoid that is strided in memory, so the owner packs it into a contiguous staging buffer (pitched device copy), NCCL
    moves that, and the receivers unpack.  Halves the bytes on the wire, but measured SLOWER than whole panels on NVLink 5 (C4 at 8
    GPUs 19.4 vs 18.6 ms, at 2 GPUs 70.4 vs 66.1 ms: the pack / unpack copies compete with the solve for HBM and SMs), so
    unified_rectrxm_pipelined keeps whole panels by default and this stays an option (`triangle_only=True`) for slower links."""
    import ctypes

    import torch
    import torch.distributed as dist

    from . import _check, load_library

    n = A_store.shape[0]
    c0, c1 = p * pc, min(n, (p + 1) * pc)
    r0, r1 = (c0, n) if uplo == "L" else (0, c1)
    w, cols, es = r1 - r0, c1 - c0, A_store.element_size()
    key = (A_store.device, A_store.dtype)
    if key not in _stage or _stage[key].numel() < n * pc:
        _stage[key] = torch.empty(n * pc, dtype=A_store.dtype, device=A_store.device)
    flat = _stage[key][:w * cols]
    lib, rank = load_library(), dist.get_rank(group)
    panel_ptr = A_store.data_ptr() + (c0 * n + r0) * es
    sp = ctypes.c_void_p(stream.cuda_stream)
    if rank == src:
        _check(lib.nla_memcpy2d_async(handle._h, flat.data_ptr(), w * es, panel_ptr, n * es, w * es, cols, 2, sp), handle._h)
    dist.broadcast(flat, src=src, group=group)
    if rank != src:
        _check(lib.nla_memcpy2d_async(handle._h, panel_ptr, n * es, flat.data_ptr(), w * es, w * es, cols, 2, sp), handle._h)


def unified_rectrxm_pipelined(side: str, uplo: str, transpose: str, alpha: float, func: str, A, B_local, src: int = 0, group=None,
                              panels: int = 8, handle=None, triangle_only: bool = False):
    """Multi-GPU call with the broadcast of A overlapped with the solve: A travels over NCCL in column
    panels, in the order the schedule consumes them, on a side stream; the solve runs on the current stream and waits for
    each panel right before the first kernel that reads it (nla_rectrxm_gated).  CUDA only."""
    import torch
    import torch.distributed as dist

    from . import panel_order, unified_rectrxm, unified_rectrxm_gated

    if not (dist.is_initialized() and dist.get_world_size(group) > 1):
        return unified_rectrxm(side, uplo, transpose, alpha, func, A, B_local, handle=handle)
    n = A.shape[0]
    if not (A.dim() == 2 and A.stride(0) == 1 and A.stride(1) == n):
        raise ValueError("pipelined broadcast needs a column-major A with leading dimension n")
    # panels that are a multiple of 1024 columns (the block-inverse order) let the Float32/Float16 solves prepare their diagonal blocks
    # panel by panel instead of waiting for all of A
    pc, npan = panel_geometry(n, panels, gran=1024 if n >= 1024 * panels else 128)
    order = panel_order(side, uplo, transpose, func, n, pc)
    dev = A.device
    if dev not in _side_streams:
        _side_streams[dev] = torch.cuda.Stream(device=dev)
    bs = _side_streams[dev]
    events = [torch.cuda.Event() for _ in range(npan)]
    bs.wait_stream(torch.cuda.current_stream(dev))   # earlier work on the caller's stream may still be using A
    from . import default_handle

    h = handle or default_handle(dev.index)
    with torch.cuda.stream(bs):
        if triangle_only:   # pack / broadcast / unpack only the trapezoid of each panel that the `uplo` triangle covers
            for p in order:
                broadcast_panel_triangle(A.t(), uplo, p, pc, src, group, bs, h)
                events[p].record(bs)
        else:
            broadcast_panels(A.t(), order, pc, src=src, group=group, on_panel=lambda p: events[p].record(bs))
    empty = B_local.shape[1] == 0 if side == "L" else B_local.shape[0] == 0
    if empty:
        torch.cuda.current_stream(dev).wait_stream(bs)
        return B_local
    return unified_rectrxm_gated(side, uplo, transpose, alpha, func, A, B_local, pc, events, handle=handle)


_host_pipe = {}


def unified_rectrxm_pipelined_host(side: str, uplo: str, transpose: str, alpha: float, func: str, A_dev, A_host, B_host, src: int = 0,
                                   group=None, panels: int = 8, handle=None):
    """End-to-end multi-GPU call with HOST buffers.  `A_host` (pinned, column-major, significant on `src` only) goes
    host -> owner GPU -> every GPU panel by panel (H2D copy + NCCL broadcast on a side stream, in consumption order) into
    `A_dev` (column-major, ld = n); the rank's right-hand sides `B_host` (column-major ndarray-like torch CPU tensor, pinned)
    are streamed through the device by the library's host pipeline (nla_rectrxm_hostb_gated: chunks of B in first-touch
    order, every launch gated on the panels of A it reads, results copied back as soon as they are final).  Synchronous."""
    import ctypes

    import torch
    import torch.distributed as dist

    from . import _ch, _check, _desc, default_handle, load_library, panel_order

    multi = dist.is_initialized() and dist.get_world_size(group) > 1
    rank = dist.get_rank(group) if multi else src
    n = A_dev.shape[0]
    dev = A_dev.device
    if dev not in _host_pipe:
        _host_pipe[dev] = torch.cuda.Stream(device=dev)
    bs = _host_pipe[dev]
    bs.wait_stream(torch.cuda.current_stream(dev))
    pc, npan = panel_geometry(n, panels)
    order = panel_order(side, uplo, transpose, func, n, pc)
    events = [torch.cuda.Event() for _ in range(npan)]
    A_store = A_dev.t()
    Ah_store = A_host.t() if rank == src else None
    h = handle or default_handle(dev.index)
    with torch.cuda.stream(bs):
        for p in order:
            rows = slice(p * pc, min(n, (p + 1) * pc))
            if rank == src:   # only the referenced triangle crosses PCIe: rows [p*pc, n) of a lower panel, [0, (p+1)*pc) of an upper one
                # (a pitched cudaMemcpy2DAsync: torch's copy_ of a non-contiguous host slice goes through a pageable staging copy
                #  and blocks -- measured 600 ms per step at 4 GPUs)
                r0, r1 = (p * pc, n) if uplo == "L" else (0, min(n, (p + 1) * pc))
                c0, c1 = rows.start, rows.stop
                es = A_dev.element_size()
                rc = load_library().nla_memcpy2d_async(h._h, A_store.data_ptr() + (c0 * n + r0) * es, n * es,
                                                       Ah_store.data_ptr() + (c0 * Ah_store.stride(0) + r0) * es, Ah_store.stride(0) * es,
                                                       (r1 - r0) * es, c1 - c0, 1, ctypes.c_void_p(bs.cuda_stream))
                _check(rc, h._h)
            if multi:
                dist.broadcast(A_store[rows], src=src, group=group)
            events[p].record(bs)
    pa, ar, ac, lda, dta = _desc(A_dev)
    if B_host.dim() != 2 or (B_host.shape[0] > 1 and B_host.stride(0) != 1):
        raise ValueError("B_host must be a column-major 2-D CPU tensor")
    m = B_host.shape[1] if side == "L" else B_host.shape[0]
    ldb = B_host.stride(1) if B_host.shape[1] > 1 else max(1, B_host.shape[0])
    evs = (ctypes.c_void_p * npan)(*[ctypes.c_void_p(e.cuda_event) for e in events])
    rc = load_library().nla_rectrxm_hostb_gated(h._h, _ch(side), _ch(uplo), _ch(transpose), _ch(func), dta, n, m, float(alpha), pa, lda,
                                                B_host.data_ptr(), ldb, pc, npan, evs)
    _check(rc, h._h)
    return B_host
