"""CPU tests of the drop-in boundary: the C-ABI library loads without a GPU, exports every symbol declared in
include/nextla_b200.h, reports errors as status codes, and its host-side schedule equals the reference's recursion."""
import ctypes
import itertools
import os
import re

import numpy as np
import pytest

from oracle import reference_port as rp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(nla):
    lib = nla.load_library()
    header = open(os.path.join(ROOT, "include", "nextla_b200.h")).read()
    declared = set(re.findall(r"\b(nla_[a-z_0-9]+)\s*\(", header))
    declared -= {"nla_context"}
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert declared == set(nla.exported_symbols())
    assert lib.nla_version() >= 100


def test_status_strings_and_invalid_handle(nla):
    lib = nla.load_library()
    for code in range(0, 11):
        assert lib.nla_status_string(code) and b"unknown" not in lib.nla_status_string(code)
    assert lib.nla_mg_destroy(None) == 8 and lib.nla_mg_sync(None) == 8 and lib.nla_mg_device_count(None) == -1
    assert lib.nla_workspace_bytes(None, b"L", b"S", 1, 1024, 1024) == -8
    assert lib.nla_destroy(None) == 8
    assert lib.nla_rectrxm(None, b"L", b"L", b"N", b"S", 0, 4, 4, 1.0, None, 4, None, 4, None) == 8
    assert lib.nla_leaf_max(0) == 1024 and lib.nla_leaf_max(7) == -1
    assert lib.nla_getrf2(None, 0, 4, 4, None, 4, None, None, None) == 8 and lib.nla_lauum(None, b"L", 0, 4, None, 4, 2, None) == 8
    assert lib.nla_laswp(None, 0, 4, 4, None, 4, 1, 4, None, 1, None) == 8


def test_create_without_gpu_fails_loudly(nla):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = ctypes.c_void_p()
    rc = nla.load_library().nla_create(ctypes.byref(h), 0)
    assert rc == 6 and not h.value  # NLA_ERR_NO_DEVICE: no CPU fallback
    with pytest.raises(nla.NextLAError):
        nla.Handle(0)
    mg = ctypes.c_void_p()
    assert nla.load_library().nla_mg_create(ctypes.byref(mg), 1, None) == 6 and not mg.value


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "nextla.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".jl")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("test oracle", "").replace("own oracle", ""), f"{f} mentions the oracle"


def _reference_trace(side, uplo, trans, func, n, threshold):
    """Run the oracle's restatement of unified_rec with recording stubs in place of the kernels and express every call
    in the normalised coordinates the library's schedule uses."""
    calls = []
    A = np.zeros((n, n), order="F")
    B = np.zeros((n, 1) if side == "L" else (1, n), order="F")
    baseA, baseB = A.ctypes.data, B.ctypes.data

    def boff(v):  # element offset of a view of B along the vector-element axis
        return (v.ctypes.data - baseB) // 8

    def leaf(Av, Bv):
        calls.append((0, boff(Bv), Av.shape[0], 0, 0))

    def mm(out, in1, in2, alpha):
        if side == "L":   # out = rows c of B, in2 = rows k of B
            calls.append((1, boff(out), out.shape[0], boff(in2), in2.shape[0]))
        else:             # out = cols c of B, in1 = cols k of B
            calls.append((1, boff(out), out.shape[1], boff(in1), in1.shape[1]))

    saved = {k: getattr(rp, k) for k in ["left_lower_trsm", "left_upper_trsm", "right_lower_trsm", "right_upper_trsm", "left_lower_trmm",
                                         "left_upper_trmm", "right_lower_trmm", "right_upper_trmm", "matmul_kernel"]}
    try:
        for k in saved:
            setattr(rp, k, mm if k == "matmul_kernel" else leaf)
        Av, u = (A.T, "U" if uplo == "L" else "L") if trans != "N" else (A, uplo)   # src/rectrxm.jl:56-59
        rp.unified_rec(func, side, u, Av, n, B, threshold)
    finally:
        for k, v in saved.items():
            setattr(rp, k, v)
    return calls


@pytest.mark.parametrize("n,leaf", [(16, 16), (64, 16), (100, 16), (300, 128), (1000, 64), (1024, 128), (777, 32)])
def test_schedule_equals_reference_recursion(nla, n, leaf):
    """The host-side schedule (flattened recursion) must issue the same leaves and updates in the same order as
    unified_rec (src/rectrxm.jl:101-198) run with threshold = leaf, for every side/uplo/trans/func."""
    import itertools

    for side, uplo, trans, func in itertools.product("LR", "LU", "NT", "SM"):
        want = _reference_trace(side, uplo, trans, func, n, leaf)
        got = [p[:5] for p in nla.plan(side, uplo, trans, func, n, leaf)]
        assert got == want, (side, uplo, trans, func)


def test_schedule_alpha_applied_exactly_once(nla):
    """alpha is folded into kernels instead of the reference's full pass over B (src/rectrxm.jl:64,72): every element
    range must be scaled by exactly one op."""
    import itertools

    for n, leaf in [(300, 128), (1024, 128), (100, 16)]:
        for side, uplo, trans, func in itertools.product("LR", "LU", "NT", "SM"):
            cover = np.zeros(n, dtype=int)
            for kind, c0, cn, k0, kn, carries in nla.plan(side, uplo, trans, func, n, leaf):
                if carries:
                    cover[c0:c0 + cn] += 1
            assert (cover == 1).all(), (n, side, uplo, trans, func)


def test_bad_arguments_return_status_codes(nla):
    lib = nla.load_library()
    assert lib.nla_plan(b"X", b"L", b"N", b"S", 8, 0, None, 0) == -1
    assert lib.nla_plan(b"L", b"L", b"N", b"S", -1, 0, None, 0) == -2
    assert lib.nla_plan(b"L", b"L", b"N", b"S", 0, 0, None, 0) == 0


def test_panel_order_follows_the_schedule(nla):
    """nla_panel_order (host-only): the column panels of A in the order the schedule first reads them -- ascending when the
    schedule walks the diagonal forward, descending otherwise; every panel exactly once; consistent with nla_plan."""
    n, pc = 4096, 512
    for side, uplo, trans, func in itertools.product("LR", "LU", "NT", "SM"):
        order = nla.panel_order(side, uplo, trans, func, n, pc)
        assert sorted(order) == list(range(n // pc)), (side, uplo, trans, func)
        # first leaf of the flattened recursion tells the direction
        first = [op for op in nla.plan(side, uplo, trans, func, n, 128) if op[0] == 0][0]
        assert order[0] == first[1] // pc
        assert order == sorted(order) or order == sorted(order, reverse=True)
    assert nla.panel_order("L", "L", "N", "S", 1000, 256) == [0, 1, 2, 3]
    assert nla.panel_order("L", "U", "N", "S", 1000, 256) == [3, 2, 1, 0]
    assert nla.panel_order("L", "L", "N", "S", 0, 256) == []
    with pytest.raises(nla.NextLAError):
        nla.panel_order("X", "L", "N", "S", 100, 10)


@pytest.mark.parametrize("n", [1024, 3000, 4096, 5000, 16384, 20000])
@pytest.mark.parametrize("cutoff", [1024, 2048])
def test_large_cutoff_schedule_invariants(nla, n, cutoff):
    """What the fused FP64 slab (cutoff 2048) and the block-inverse Float32/Float16 leaves (cutoff 1024) rely on, for every variant
    and ragged orders: (1) the schedule is still the reference recursion (src/rectrxm.jl:129-134 split rule) with that threshold;
    (2) every leaf starts at a multiple of the cutoff and only the last block along the diagonal is ragged (the inverse workspace
    is indexed by block); (3) in a solve, a leaf that follows an update lies inside that update's output range (the update's epilogue writes
    the leaf's copy of V, GemmTcParams::dup); (4) the ops of a solve partition the flops n^2 exactly."""
    for side, uplo, trans, func in itertools.product("LR", "LU", "NT", "SM"):
        ops = nla.plan(side, uplo, trans, func, n, cutoff)
        assert [p[:5] for p in ops] == _reference_trace(side, uplo, trans, func, n, cutoff)
        leaves = sorted((c0, cn) for kind, c0, cn, k0, kn, _ in ops if kind == 0)
        assert [c0 for c0, _ in leaves] == list(range(0, n, cutoff))
        assert all(cn == cutoff for _, cn in leaves[:-1]) and leaves[-1][1] == n - leaves[-1][0]
        for prev, cur in zip(ops, ops[1:]):
            if func == "S" and cur[0] == 0 and prev[0] == 1:   # (multiplies have no block-inverse leaves)
                assert prev[1] <= cur[1] and cur[1] + cur[2] <= prev[1] + prev[2], (side, uplo, trans, func, prev, cur)
        flops = sum(cn * cn if kind == 0 else 2 * cn * kn for kind, c0, cn, k0, kn, _ in ops)
        assert flops == n * n


@pytest.mark.parametrize("n,cutoff,slabs,resident", [(4096, 1024, 1, False), (5000, 1024, 4, False), (16384, 1024, 4, False), (3000, 128, 3, True), (9000, 2048, 2, False)])
def test_host_pipeline_transfer_plan(nla, n, cutoff, slabs, resident):
    """The plan behind nla_rectrxm_host (host-only, nla_host_plan), for every side/uplo/trans/func: (1) exactly the 1024 x 1024 tiles of A that
    touch the referenced triangle are uploaded, once each (none when A is resident), and every (chunk, slab) of B once; (2) no op runs
    before everything it reads has been queued: its A block and its rows of B -- the rows it writes and, for an update, the rows it
    multiplies -- have transfer indices <= need[op][slab]; (3) the pieces of a cut update tile its output range; (4) a chunk is downloaded
    after the last op that writes it and no later op writes it; (5) transfers are issued in first-touch order (need is non-decreasing
    along the schedule for a slab, and slab q never waits for more than slab q+1 at the same op)."""
    TS = 1024
    nt = -(-n // TS)
    for side, uplo, trans, func in itertools.product("LR", "LU", "NT", "SM"):
        ops, xfers, need, last = nla.host_plan(side, uplo, trans, func, n, cutoff, slabs, resident)
        right, tr = side == "R", trans != "N"
        teff_trans = tr != right
        lower = uplo == "L"
        a_idx = {(i, j): x for x, (k, i, j) in enumerate(xfers) if k == 0}
        b_idx = {(i, j): x for x, (k, i, j) in enumerate(xfers) if k == 1}
        assert len(a_idx) + len(b_idx) == len(xfers)                                   # (1) nothing twice
        want_tiles = set() if resident else {(i, j) for i in range(nt) for j in range(nt) if (i >= j if lower else i <= j)}
        assert set(a_idx) == want_tiles
        assert set(b_idx) == {(c, q) for c in range(nt) for q in range(slabs)}
        # (3) the schedule is the plain one with the wide updates cut at multiples of 1024
        plain = nla.plan(side, uplo, trans, func, n, cutoff)
        rebuilt, i = [], 0
        while i < len(ops):
            kind, c0, cn, k0, kn, _ = ops[i]
            if kind == 1:
                j = i
                while j + 1 < len(ops) and ops[j + 1][0] == 1 and ops[j + 1][3:5] == (k0, kn) and ops[j + 1][1] == ops[j][1] + ops[j][2] and ops[j][2] <= TS and (ops[j][1] + ops[j][2]) % TS == 0:
                    j += 1
                rebuilt.append((1, c0, ops[j][1] + ops[j][2] - c0, k0, kn))
                i = j + 1
            else:
                rebuilt.append((0, c0, cn, 0, 0)); i += 1
        assert rebuilt == [p[:5] for p in plain], (side, uplo, trans, func)
        writers = {}
        for oi, (kind, c0, cn, k0, kn, _) in enumerate(ops):
            rows_b = [(c0, c0 + cn)] + ([(k0, k0 + kn)] if kind == 1 else [])
            if kind == 0:
                ar, ac = (c0, c0 + cn), (c0, c0 + cn)
            else:
                ar, ac = ((k0, k0 + kn), (c0, c0 + cn)) if teff_trans else ((c0, c0 + cn), (k0, k0 + kn))
            for q in range(slabs):
                nd = need[oi][q]
                for lo, hi in rows_b:                                                     # (2) B rows
                    for c in range(lo // TS, (hi - 1) // TS + 1):
                        assert b_idx[(c, q)] <= nd
                if not resident:                                                          # (2) A block (tiles inside the triangle)
                    for ti in range(ar[0] // TS, (ar[1] - 1) // TS + 1):
                        for tj in range(ac[0] // TS, (ac[1] - 1) // TS + 1):
                            if (ti, tj) in a_idx:
                                assert a_idx[(ti, tj)] <= nd
                if oi > 0:
                    assert need[oi][q] >= -1 and max(need[oi][q], need[oi - 1][q]) >= need[oi - 1][q]
                if q + 1 < slabs:
                    assert need[oi][q] <= need[oi][q + 1]                                 # (5) slab order within an op
            for c in range(c0 // TS, (c0 + cn - 1) // TS + 1):
                writers[c] = oi
        assert [writers[c] for c in range(nt)] == last                                    # (4)


def test_julia_binding_matches_the_c_prototypes(nla):
    """The Julia side of the drop-in (nextla.jl_b200/julia/NextLAB200.jl) cannot be executed here (no Julia in the image), so it is
    checked statically: every `ccall` names a symbol that include/nextla_b200.h declares and the library exports, and its argument-type
    tuple has exactly as many entries as the C prototype has parameters; every wrapper that takes matrices validates their shapes
    (a mismatched call must raise DimensionMismatch instead of reaching the device)."""
    lib = nla.load_library()
    header = open(os.path.join(ROOT, "include", "nextla_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(nla_[a-z_0-9]+)\s*\(([^;{]*?)\)\s*;", header):
        args = m.group(2).strip()
        protos[m.group(1)] = 0 if args in ("", "void") else args.count(",") + 1
    jl = open(os.path.join(ROOT, "nextla.jl_b200", "julia", "NextLAB200.jl")).read()
    calls = re.findall(r"ccall\(\((?::(\w+)|\$\(QuoteNode\(cfun\)\)),\s*libnextla\),\s*\w+,\s*\(([^)]*)\)", jl)
    assert len(calls) >= 15
    for name, types in calls:
        names = [name] if name else ["nla_trsm_leaf", "nla_trmm_leaf"]
        ntypes = len([t for t in types.split(",") if t.strip()])
        for nm in names:
            assert nm in protos, f"{nm} is not declared in the header"
            assert hasattr(lib, nm), f"{nm} is not exported"
            assert ntypes == protos[nm], f"ccall of {nm} passes {ntypes} argument types, the C prototype has {protos[nm]}"
    # shape validation in every matrix-taking wrapper; handles keyed by (device, stream)
    assert jl.count("check_shapes(") >= 5 and "DimensionMismatch" in jl
    assert "Dict{Tuple{Int,UInt},Ptr{Cvoid}}" in jl and "CUDA.stream().handle" in jl
