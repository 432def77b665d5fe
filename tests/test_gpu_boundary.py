"""GPU tests of the boundary's resource behaviour (include/nextla_b200.h "Workspace control", handle / stream rules): what a Julia host
relies on when it calls the library from several tasks -- none of it is arithmetic, all of it goes through the C ABI."""
import itertools

import numpy as np
import pytest

from oracle import reference_port as rp

pytestmark = pytest.mark.gpu

TOL = {np.float32: 1e-5, np.float16: 1e-2}
CODE = {np.float64: 0, np.float32: 1, np.float16: 2}


def _run(nla, h, side, uplo, trans, alpha, func, A, B0, stream=None):
    import torch

    dA, dB = nla.colmajor(A), nla.colmajor(B0)
    nla.unified_rectrxm(side, uplo, trans, alpha, func, dA, dB, handle=h, stream=stream)
    torch.cuda.synchronize()
    return nla.to_numpy(dB)


@pytest.mark.parametrize("dtype", [np.float32, np.float16])
def test_warm_handle_allocates_nothing_and_reserve_presizes(nla, dtype):
    """First call grows the workspaces with the stream-ordered allocator; a second call of the same shape must not allocate at all;
    after nla_reserve even the first call of a shape allocates nothing."""
    h = nla.Handle(0)
    try:
        n, m = 2304, 520
        for side, func in itertools.product("LR", "SM"):
            A, B0 = rp.make_inputs(n, m, side, "L", dtype, seed=3, recipe="scaled")
            _run(nla, h, side, "L", "N", 1.0, func, A, B0)
        warm = h.get_option("ws_allocs")
        assert warm > 0
        for side, func in itertools.product("LR", "SM"):
            A, B0 = rp.make_inputs(n, m, side, "L", dtype, seed=3, recipe="scaled")
            got = _run(nla, h, side, "L", "N", 1.0, func, A, B0)
            assert rp.error_metric(side, "L", "N", 1.0, func, A, B0, got) < TOL[dtype]
        assert h.get_option("ws_allocs") == warm
        # a larger shape, reserved ahead of the call
        n2, m2 = 4096, 1030
        for side, func in itertools.product("LR", "SM"):
            assert h.workspace_bytes(side, func, CODE[dtype], n2, m2) > 0
            h.reserve(side, func, CODE[dtype], n2, m2)
        reserved = h.get_option("ws_allocs")
        for side, func in itertools.product("LR", "SM"):
            A, B0 = rp.make_inputs(n2, m2, side, "U", dtype, seed=4, recipe="scaled")
            got = _run(nla, h, side, "U", "T", 1.0, func, A, B0)
            assert rp.error_metric(side, "U", "T", 1.0, func, A, B0, got) < TOL[dtype]
        assert h.get_option("ws_allocs") == reserved
    finally:
        h.close()


@pytest.mark.parametrize("dtype", [np.float64, np.float32, np.float16])
def test_caller_provided_workspace(nla, dtype):
    """nla_set_workspace: with an arena of nla_workspace_bytes the library allocates nothing and the result is bit-identical to the
    library-owned run; with a smaller arena the call degrades (128-wide leaves / in-place multiply / native right side) and stays
    within tolerance; an arena that is too small even for that is NLA_ERR_WORKSPACE (Float32/Float16 only: Float64 always has the
    workspace-free native schedule)."""
    import torch

    n, m = 2304, 520
    h = nla.Handle(0)
    ref = nla.Handle(0)
    try:
        for side, func in itertools.product("LR", "SM"):
            A, B0 = rp.make_inputs(n, m, side, "L", dtype, seed=8, recipe="scaled")
            want = _run(nla, ref, side, "L", "N", 1.5, func, A, B0)
            need = h.workspace_bytes(side, func, CODE[dtype], n, m)
            if dtype == np.float64:
                assert (need > 0) == (side == "R")
            else:
                assert need > 0
            arena = torch.empty(max(need, 256), dtype=torch.uint8, device="cuda")
            h.set_workspace(arena)
            got = _run(nla, h, side, "L", "N", 1.5, func, A, B0)
            assert h.get_option("ws_allocs") == 0
            assert np.array_equal(got, want), (side, func)
            if need > 0:
                small = torch.empty(max(256, need // 3), dtype=torch.uint8, device="cuda")
                h.set_workspace(small)
                if dtype == np.float64:
                    got = _run(nla, h, side, "L", "N", 1.5, func, A, B0)
                    assert rp.error_metric(side, "L", "N", 1.5, func, A, B0, got) < 1e-13
                else:
                    diag_only = n * 128 * np.dtype(dtype).itemsize + 4096
                    small = torch.empty(diag_only, dtype=torch.uint8, device="cuda")
                    h.set_workspace(small)
                    got = _run(nla, h, side, "L", "N", 1.5, func, A, B0)
                    assert rp.error_metric(side, "L", "N", 1.5, func, A, B0, got) < TOL[dtype], (side, func)
                    h.set_workspace(torch.empty(1024, dtype=torch.uint8, device="cuda"))
                    with pytest.raises(nla.NextLAError, match="status 9"):
                        _run(nla, h, side, "L", "N", 1.5, func, A, B0)
                assert h.get_option("ws_allocs") == 0
            h.set_workspace(None)
    finally:
        h.close()
        ref.close()


def test_one_handle_per_stream_concurrent_calls(nla):
    """Two streams on one GPU, each with ITS OWN handle (what the bindings' (device, stream)-keyed cache hands out), running Float16
    solves that use the per-handle workspaces at the same time: both results must equal the serial ones."""
    import torch

    n, m = 3072, 2048
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    assert nla.default_handle(0, s1) is not nla.default_handle(0, s2)
    assert nla.default_handle(0, s1) is nla.default_handle(0, s1)
    cases = [("L", "L", "N", np.float16), ("R", "U", "T", np.float16), ("L", "U", "N", np.float32), ("R", "L", "N", np.float32)]
    ins = [rp.make_inputs(n, m, sd, up, dt, seed=20 + i, recipe="scaled") for i, (sd, up, tr, dt) in enumerate(cases)]
    serial = [_run(nla, None, sd, up, tr, 1.0, "S", A, B0) for (sd, up, tr, dt), (A, B0) in zip(cases, ins)]
    dev = [(nla.colmajor(A), nla.colmajor(B0)) for A, B0 in ins]
    torch.cuda.synchronize()
    for rep in range(3):
        for i, ((sd, up, tr, dt), (dA, dB)) in enumerate(zip(cases, dev)):
            dB.copy_(nla.colmajor(ins[i][1]))
        torch.cuda.synchronize()
        for i, ((sd, up, tr, dt), (dA, dB)) in enumerate(zip(cases, dev)):
            st = s1 if i % 2 == 0 else s2
            with torch.cuda.stream(st):
                nla.unified_rectrxm(sd, up, tr, 1.0, "S", dA, dB, stream=st)
        torch.cuda.synchronize()
        for i, (dA, dB) in enumerate(dev):
            assert np.array_equal(nla.to_numpy(dB), serial[i]), (rep, cases[i][:3])


def test_wrong_device_handle_is_rejected(nla, gpu):
    import torch

    if torch.cuda.device_count() < 2:
        A = nla.colmajor(np.eye(8)); B = nla.colmajor(np.ones((8, 2)))
        nla.unified_rectrxm("L", "L", "N", 1.0, "S", A, B, handle=gpu)   # same device: fine
        return
    h1 = nla.Handle(1)
    try:
        A = nla.colmajor(np.eye(8)); B = nla.colmajor(np.ones((8, 2)))   # on cuda:0
        with pytest.raises(nla.NextLAError):
            nla.unified_rectrxm("L", "L", "N", 1.0, "S", A, B, handle=h1)
        # a call on device 1 must leave the caller's current device (0) untouched
        A1 = A.to("cuda:1").t().contiguous().t(); B1 = B.to("cuda:1").t().contiguous().t()
        assert torch.cuda.current_device() == 0
        nla.unified_rectrxm("L", "L", "N", 2.0, "S", A1, B1, handle=h1)
        assert torch.cuda.current_device() == 0
        torch.cuda.synchronize(1)
        assert np.allclose(nla.to_numpy(B1), 2.0)
    finally:
        h1.close()


@pytest.mark.parametrize("dtype", [np.float64, np.float16])
def test_single_process_multi_gpu_entry(nla, gpu, dtype):
    """nla_mg_* (SURVEY.md 8(b)): the library-owned multi-GPU pipeline -- NCCL communicator, per-GPU streams, panel broadcast of A,
    gated solves -- driven from one host thread through ctypes, no torch.distributed.  Uses every GPU of the box (1 here on a
    one-GPU box: same code path minus the NCCL calls; `gpurun --gpus 2` runs it with a real broadcast).  Device-resident and
    host-buffer variants against the single-GPU call, ragged shards, an empty shard, a non-zero root."""
    import torch

    ng = torch.cuda.device_count()
    tol = 1e-13 if dtype == np.float64 else 1e-2
    mg = nla.MultiGPU(ngpu=ng)
    try:
        n = 2304
        per = [300 + 136 * i for i in range(ng)]
        if ng > 1:
            per[-1] = 0                                    # an empty shard must be harmless
        root = ng - 1
        for side, uplo, trans, func in [("L", "L", "N", "S"), ("R", "U", "T", "S"), ("L", "U", "N", "M"), ("R", "L", "N", "M")]:
            m = sum(per)
            A, B0 = rp.make_inputs(n, m, side, uplo, dtype, seed=50, recipe="scaled")
            want = _run(nla, None, side, uplo, trans, 1.25, func, A, B0)
            offs = np.cumsum([0] + per)
            parts = [np.asfortranarray(B0[:, offs[i]:offs[i + 1]] if side == "L" else B0[offs[i]:offs[i + 1], :]) for i in range(ng)]
            # device-resident
            dA = nla.colmajor(A, device=f"cuda:{root}")
            shards = [nla.colmajor(parts[i], device=f"cuda:{i}") for i in range(ng)]
            for d in range(ng):
                torch.cuda.synchronize(d)
            mg.rectrxm(side, uplo, trans, 1.25, func, dA, root, shards)
            mg.sync()
            got = np.concatenate([nla.to_numpy(s) for s in shards], axis=1 if side == "L" else 0)
            assert np.array_equal(got, want) or rp.error_metric(side, uplo, trans, 1.25, func, A, B0, got) < tol, (side, uplo, trans, func)
            assert rp.error_metric(side, uplo, trans, 1.25, func, A, B0, got) < tol
            # host buffers
            hparts = [p.copy(order="F") for p in parts]
            mg.rectrxm_host(side, uplo, trans, 1.25, func, A, hparts)
            got = np.concatenate(hparts, axis=1 if side == "L" else 0)
            assert rp.error_metric(side, uplo, trans, 1.25, func, A, B0, got) < tol, (side, uplo, trans, func)
        assert torch.cuda.current_device() == 0
    finally:
        mg.close()


@pytest.mark.parametrize("uplo,trans", [("L", "N"), ("U", "N"), ("L", "T"), ("U", "T")])
def test_streaming_host_pipeline(nla, gpu, uplo, trans):
    """nla_rectrxm_host, Float64 left-side solve with enough right-hand sides to fill the machine: ONE streaming launch of the
    row-split slab kernel (operands arrive chunk by chunk behind device flags, finished chunks are downloaded while the kernel
    runs).  Forward and backward walks of the diagonal, stored / transposed A, ragged order (not a multiple of 128), alpha != 1,
    pinned and pageable buffers; result against OpenBLAS and bit-identical to the same kernel on device-resident data."""
    import torch

    n, m = 2184, 7168
    A, B0 = rp.make_inputs(n, m, "L", uplo, np.float64, seed=77, recipe="scaled")
    want = rp.blas_reference("L", uplo, trans, 1.25, "S", A, B0)
    gpu.launch_count(reset=True)
    B = B0.copy(order="F")
    nla.unified_rectrxm_host("L", uplo, trans, 1.25, "S", A, B, handle=gpu)
    assert gpu.launch_count() == 1                                   # the whole solve is one kernel
    assert np.linalg.norm(B - want) / np.linalg.norm(want) < 1e-13
    assert rp.error_metric("L", uplo, trans, 1.25, "S", A, B0, B) < 1e-14
    # pinned buffers, and the same kernel on resident data (macro >= n: one fused-slab launch)
    hA = torch.from_numpy(np.ascontiguousarray(A.T)).pin_memory()
    hB = torch.from_numpy(np.ascontiguousarray(B0.T)).pin_memory()
    rc = nla.load_library().nla_rectrxm_host(gpu._h, b"L", uplo.encode(), trans.encode(), b"S", 0, n, m, 1.25, hA.data_ptr(), n, hB.data_ptr(), n)
    assert rc == 0
    assert np.array_equal(np.asfortranarray(hB.numpy().T), B)
    gpu.set_option("macro", 4096); gpu.set_option("streams", 1)
    try:
        dA, dB = nla.colmajor(A), nla.colmajor(B0)
        nla.unified_rectrxm("L", uplo, trans, 1.25, "S", dA, dB, handle=gpu)
        torch.cuda.synchronize()
        assert np.array_equal(nla.to_numpy(dB), B)
    finally:
        gpu.set_option("macro", -1); gpu.set_option("streams", 0)
    # the chunked pipeline (option host_stream = 0) must agree to rounding
    gpu.set_option("host_stream", 0)
    try:
        B2 = B0.copy(order="F")
        nla.unified_rectrxm_host("L", uplo, trans, 1.25, "S", A, B2, handle=gpu)
    finally:
        gpu.set_option("host_stream", 1)
    assert np.linalg.norm(B2 - B) / np.linalg.norm(B) < 1e-13


@pytest.mark.parametrize("uplo,trans", [("L", "N"), ("U", "N"), ("L", "T"), ("U", "T")])
def test_gated_one_launch_solve(nla, gpu, uplo, trans):
    """nla_rectrxm_gated, Float64 left-side solve with enough right-hand sides: A arrives in column panels (here copied on a side
    stream, each panel behind a spin kernel, the rest of A NaN until then) and the solve is still ONE launch of the row-split slab
    kernel -- a side stream counts the panel events into a device word, every block row waits for the panels that hold its part of A.
    Result bit-identical to the plain call."""
    import torch

    n, m, pc = 2048, 7168, 256
    A, B0 = rp.make_inputs(n, m, "L", uplo, np.float64, seed=61, recipe="scaled")
    dA_full = nla.colmajor(A)
    dB = nla.colmajor(B0)
    nla.unified_rectrxm("L", uplo, trans, 1.25, "S", dA_full, dB, handle=gpu)
    torch.cuda.synchronize()
    want = nla.to_numpy(dB)
    assert rp.error_metric("L", uplo, trans, 1.25, "S", A, B0, want) < 1e-14
    dA = torch.full_like(dA_full, float("nan"))
    dB = nla.colmajor(B0)
    order = nla.panel_order("L", uplo, trans, "S", n, pc)
    side_stream = torch.cuda.Stream()
    events = [torch.cuda.Event() for _ in range(n // pc)]
    torch.cuda.synchronize()
    with torch.cuda.stream(side_stream):
        torch.cuda._sleep(40_000_000)      # ~20 ms before the first panel: the kernel is resident and really has to wait
        for p in order:
            dA[:, p * pc:(p + 1) * pc].copy_(dA_full[:, p * pc:(p + 1) * pc])
            events[p].record(side_stream)
    gpu.set_option("gated_stream", 1)   # opt-in: safe here, the panels arrive through the copy engine
    try:
        gpu.launch_count(reset=True)
        nla.unified_rectrxm_gated("L", uplo, trans, 1.25, "S", dA, dB, pc, events, handle=gpu)
        assert gpu.launch_count() == 1
        torch.cuda.synchronize()
    finally:
        gpu.set_option("gated_stream", 0)
    got = nla.to_numpy(dB)
    assert np.isfinite(got).all()
    assert np.array_equal(got, want), (uplo, trans)
