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
_upload_streams = {}


def gpu_numa_node(local_rank: int):
    """NUMA node of a GPU from sysfs (None when it cannot be determined, e.g. single-socket boxes report -1)."""
    import torch

    try:
        pr = torch.cuda.get_device_properties(local_rank)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip())
        return node if node >= 0 else None
    except Exception:  # noqa: BLE001
        return None


def _node_cpus(node: int):
    cpus = set()
    for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
        if "-" in part:
            a, b = part.split("-")
            cpus.update(range(int(a), int(b) + 1))
        elif part:
            cpus.add(int(part))
    return cpus


_numa_state = {}


def bind_numa_local(local_rank: int) -> bool:
    """Pin this process to the CPUs of the NUMA node its GPU hangs off, so that the pinned host buffers it allocates afterwards
    (first touch by this process) live in the memory next to that GPU's PCIe root and host<->device copies do not cross the
    socket interconnect.  One process per GPU: each rank binds itself.  Returns False when the topology is not exposed."""
    import os

    node = gpu_numa_node(local_rank)
    info = {"gpu_numa_node": node, "bound": False}
    try:
        if node is not None:
            allowed = os.sched_getaffinity(0)
            cpus = _node_cpus(node) & allowed
            if cpus:
                os.sched_setaffinity(0, cpus)
                info.update(bound=True, cpus=len(cpus))
    except Exception as e:  # noqa: BLE001
        info["error"] = str(e)
    _numa_state[local_rank] = info
    return info["bound"]


def numa_report(local_rank: int):
    return _numa_state.get(local_rank, {"gpu_numa_node": gpu_numa_node(local_rank), "bound": False})


class HostSharedMatrix:
    """An n x n matrix in host memory that EVERY rank of the node can DMA from: a POSIX shared-memory file mapped by all ranks and
    page-locked by each (cudaHostRegister).  `tensor` is the contiguous (n x n) storage (= the transpose view of the column-major
    matrix, like A.t() of a device matrix).  Lets each rank upload its share of the panels of A through its own PCIe link."""

    def __init__(self, n: int, dtype, rank: int, local_rank: int, world: int, name: str = "nla_A"):
        import os

        import numpy as np
        import torch
        import torch.distributed as dist

        self.path = f"/dev/shm/{name}_{os.environ.get('MASTER_PORT', '0')}_{os.getuid()}"
        self.rank, self.world = rank, world
        npdt = {torch.float64: np.float64, torch.float32: np.float32, torch.float16: np.float16}[dtype]
        nbytes = n * n * np.dtype(npdt).itemsize
        if rank == 0:
            with open(self.path, "wb") as f:
                f.truncate(nbytes)
        if world > 1:
            dist.barrier()
        self.mm = np.memmap(self.path, dtype=npdt, mode="r+", shape=(n, n))
        self.tensor = torch.from_numpy(self.mm)
        self.registered = False
        rc = torch.cuda.cudart().cudaHostRegister(self.tensor.data_ptr(), nbytes, 0)
        self.registered = int(rc) == 0
        if world > 1:
            dist.barrier()
        if rank == 0:
            try:
                os.unlink(self.path)     # the mappings keep the memory alive; nothing is left behind in /dev/shm
            except OSError:
                pass

    def close(self):
        import torch

        if self.registered:
            torch.cuda.cudart().cudaHostUnregister(self.tensor.data_ptr())
            self.registered = False


def panel_roots(order, world: int, src):
    """Which rank uploads / broadcasts each panel: `src` when given (A lives in that rank's memory only), otherwise round-robin in
    consumption order (A in host memory every rank can read: rank k uploads the k-th, (k+N)-th, ... consumed panel)."""
    if src is not None:
        return {p: src for p in order}
    return {p: i % world for i, p in enumerate(order)}


def unified_rectrxm_pipelined_host(side: str, uplo: str, transpose: str, alpha: float, func: str, A_dev, A_host, B_host, src=0,
                                   group=None, panels: int = 8, handle=None):
    """End-to-end multi-GPU call with HOST buffers.  `A_host` (pinned, column-major) goes host -> GPU -> every GPU panel by panel
    (H2D copy of the referenced trapezoid + NCCL broadcast, in consumption order) into `A_dev` (column-major, ld = n); the rank's
    right-hand sides `B_host` (column-major torch CPU tensor, pinned) are streamed through the device by the library's host pipeline
    (nla_rectrxm_hostb_gated: chunks of B in first-touch order, every launch gated on the panels of A it reads, results copied back as
    soon as they are final).  Synchronous.
    src = r: A_host is significant on rank r only; r uploads and broadcasts every panel.
    src = None: A_host is readable by every rank (HostSharedMatrix); the uploads are spread round-robin over the ranks -- every PCIe
    link carries 1/N of A next to its own rank's B -- and each rank is the NCCL root of the panels it uploaded."""
    import ctypes

    import torch
    import torch.distributed as dist

    from . import _ch, _check, _desc, default_handle, load_library, panel_order

    multi = dist.is_initialized() and dist.get_world_size(group) > 1
    world = dist.get_world_size(group) if multi else 1
    rank = dist.get_rank(group) if multi else (src if src is not None else 0)
    n = A_dev.shape[0]
    dev = A_dev.device
    if dev not in _host_pipe:
        _host_pipe[dev] = torch.cuda.Stream(device=dev)
        _upload_streams[dev] = torch.cuda.Stream(device=dev)
    bs, us = _host_pipe[dev], _upload_streams[dev]
    cur = torch.cuda.current_stream(dev)
    bs.wait_stream(cur)
    us.wait_stream(cur)
    pc, npan = panel_geometry(n, panels)
    order = panel_order(side, uplo, transpose, func, n, pc)
    roots = panel_roots(order, world, src)
    events = [torch.cuda.Event() for _ in range(npan)]
    A_store = A_dev.t()
    h = handle or default_handle(dev.index)
    lib = load_library()
    es = A_dev.element_size()
    # uploads first, all of this rank's panels in consumption order on their own stream: they start at once on every rank and are
    # not serialised behind the broadcasts of other ranks' panels
    up_events = {}
    for p in order:
        if roots[p] != rank:
            continue
        # only the referenced triangle crosses PCIe: rows [p*pc, n) of a lower panel, [0, (p+1)*pc) of an upper one
        # (a pitched cudaMemcpy2DAsync: torch's copy_ of a non-contiguous host slice goes through a pageable staging copy and blocks)
        c0, c1 = p * pc, min(n, (p + 1) * pc)
        r0, r1 = (c0, n) if uplo == "L" else (0, c1)
        Ah_store = A_host.t()
        rc = lib.nla_memcpy2d_async(h._h, A_store.data_ptr() + (c0 * n + r0) * es, n * es,
                                    Ah_store.data_ptr() + (c0 * Ah_store.stride(0) + r0) * es, Ah_store.stride(0) * es,
                                    (r1 - r0) * es, c1 - c0, 1, ctypes.c_void_p(us.cuda_stream))
        _check(rc, h._h)
        up_events[p] = torch.cuda.Event()
        up_events[p].record(us)
    with torch.cuda.stream(bs):
        for p in order:
            if p in up_events:
                bs.wait_event(up_events[p])
            if multi:
                dist.broadcast(A_store[p * pc:min(n, (p + 1) * pc)], src=roots[p], group=group)
            events[p].record(bs)
    pa, ar, ac, lda, dta = _desc(A_dev)
    if B_host.dim() != 2 or (B_host.shape[0] > 1 and B_host.stride(0) != 1):
        raise ValueError("B_host must be a column-major 2-D CPU tensor")
    m = B_host.shape[1] if side == "L" else B_host.shape[0]
    ldb = B_host.stride(1) if B_host.shape[1] > 1 else max(1, B_host.shape[0])
    evs = (ctypes.c_void_p * npan)(*[ctypes.c_void_p(e.cuda_event) for e in events])
    rc = lib.nla_rectrxm_hostb_gated(h._h, _ch(side), _ch(uplo), _ch(transpose), _ch(func), dta, n, m, float(alpha), pa, lda,
                                     B_host.data_ptr(), ldb, pc, npan, evs)
    _check(rc, h._h)
    return B_host
