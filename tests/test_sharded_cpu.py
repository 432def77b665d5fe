"""world_size-2 gloo tests (CPU) of the N > 1 host logic: shard ranges, the broadcast of A and the per-rank solve.
The CUDA solver cannot run here, so the test injects the oracle as the per-rank solver: what is being tested is the
sharding / collective plumbing of nextla.jl_b200/sharded.py, not the arithmetic."""
import os
import socket

import numpy as np
import pytest


def test_shard_range_partitions(nla):
    from importlib import import_module

    sh = import_module(nla.__name__ + ".sharded")
    for m in [0, 1, 100, 128, 129, 16384, 65536, 1000003]:
        for world in [1, 2, 3, 4, 8]:
            cover = 0
            for r in range(world):
                v0, nv = sh.shard_range(m, world, r)
                assert v0 == min(m, cover) and nv >= 0
                if nv and v0 + nv < m:
                    assert nv % 128 == 0
                cover += nv
            assert cover == m
    with pytest.raises(ValueError):
        sh.shard_range(10, 2, 2)


def _worker(rank, world, port, side, func, n, m, out_dir):
    import sys

    import torch
    import torch.distributed as dist

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import __graft_entry__ as ge
    from importlib import import_module

    from oracle import c_port
    from oracle import reference_port as rp

    nla = ge.load_package()
    sh = import_module(nla.__name__ + ".sharded")
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    A_np, B_np = rp.make_inputs(n, m, side, "L", np.float64, seed=5)
    A = torch.from_numpy(np.ascontiguousarray(A_np.T)).t() if rank == 0 else torch.zeros(n, n, dtype=torch.float64).t()
    v0, nv = sh.shard_range(m, world, rank, gran=8)
    Bl = np.asfortranarray(B_np[:, v0:v0 + nv] if side == "L" else B_np[v0:v0 + nv, :])
    Bt = torch.from_numpy(np.ascontiguousarray(Bl.T)).t()

    def solver(side, uplo, trans, alpha, func, A_t, B_t):
        a = np.asfortranarray(A_t.numpy())
        b = np.asfortranarray(B_t.numpy())
        c_port.unified_rectrxm(side, uplo, trans, alpha, func, a, b)
        B_t.copy_(torch.from_numpy(b))

    sh.unified_rectrxm_sharded(side, "L", "N", 1.5, func, A, Bt, src=0, solver=solver)
    np.save(os.path.join(out_dir, f"shard{rank}.npy"), Bt.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("side,func", [("L", "S"), ("R", "M")])
def test_two_rank_gloo_sharded_solve(nla, tmp_path, side, func):
    import torch.multiprocessing as mp

    from oracle import c_port
    from oracle import reference_port as rp

    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    n, m, world = 96, 20, 2
    mp.spawn(_worker, args=(world, port, side, func, n, m, str(tmp_path)), nprocs=world, join=True)
    A, B0 = rp.make_inputs(n, m, side, "L", np.float64, seed=5)
    want = c_port.unified_rectrxm(side, "L", "N", 1.5, func, A, B0.copy(order="F"))
    parts = [np.load(tmp_path / f"shard{r}.npy") for r in range(world)]
    got = np.concatenate(parts, axis=1 if side == "L" else 0)
    assert np.array_equal(got, want)  # sharding RHS vectors cannot change any bit of the per-vector arithmetic


def _panel_worker(rank, world, port, n, pc, order, out_dir):
    import sys

    import torch
    import torch.distributed as dist

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import __graft_entry__ as ge
    from importlib import import_module

    nla = ge.load_package()
    sh = import_module(nla.__name__ + ".sharded")
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(3)
    full = torch.rand(n, n, dtype=torch.float64, generator=g)
    A = full.clone().t() if rank == 0 else torch.zeros(n, n, dtype=torch.float64).t()   # column-major n x n
    seen = []
    snapshots = []

    def on_panel(p):
        seen.append(p)
        snapshots.append(A[:, p * pc:min(n, (p + 1) * pc)].clone())

    sh.broadcast_panels(A.t(), order, pc, src=0, on_panel=on_panel)
    ok = seen == list(order) and torch.equal(A, full.t())
    for p, snap in zip(seen, snapshots):   # a panel is complete when its callback runs (what the CUDA path records an event for)
        ok = ok and torch.equal(snap, full.t()[:, p * pc:min(n, (p + 1) * pc)])
    np.save(os.path.join(out_dir, f"ok{rank}.npy"), np.array([ok]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_panel_broadcast(nla, tmp_path):
    """The pipelined broadcast of A (sharded.broadcast_panels): column panels in the schedule's consumption order
    (nla_panel_order), ragged last panel, every rank ends with the owner's matrix."""
    import torch.multiprocessing as mp
    from importlib import import_module

    sh = import_module(nla.__name__ + ".sharded")
    n = 200
    pc, npan = sh.panel_geometry(n, 3, gran=8)
    assert pc % 8 == 0 and (npan - 1) * pc < n <= npan * pc
    order = nla.panel_order("L", "U", "N", "S", n, pc)   # backward walk: last panel first
    assert order == list(range(npan - 1, -1, -1))
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_panel_worker, args=(2, port, n, pc, order, str(tmp_path)), nprocs=2, join=True)
    assert all(bool(np.load(tmp_path / f"ok{r}.npy")[0]) for r in range(2))
