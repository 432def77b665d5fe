# Probe: BASELINE configs 4 and 5 (and any other shape) on N GPUs, one process per GPU (torchrun), RHS sharded by rank,
# A broadcast from rank 0 inside every timed step (panel-pipelined, sharded.unified_rectrxm_pipelined).
#   torchrun --nproc-per-node 2 probes/bench_configs.py --config C4      FP16 right/lower TRSM n=32768, 131072 RHS rows in total
#   torchrun --nproc-per-node 2 probes/bench_configs.py --config C5      FP64 left/lower TRMM n=32768, 8192 RHS per GPU (weak scaling)
# Prints one JSON line on rank 0: whole-job TFLOP/s (max-over-ranks CUDA-event time), backward error of a slab of 256 vectors.
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import __graft_entry__ as ge
from importlib import import_module

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="C4"); ap.add_argument("--steps", type=int, default=3); ap.add_argument("--warmup", type=int, default=2)
ap.add_argument("--no-pipeline", action="store_true"); ap.add_argument("--triangle", type=int, default=-1, help="1/0: broadcast only the referenced triangle of each panel (default: library default)")
a = ap.parse_args()
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
nla = ge.load_package(); sh = import_module(nla.__name__ + ".sharded"); h = nla.Handle(local)
if a.config == "C4":
    dt, side, uplo, trans, func, n = torch.float16, "R", "L", "N", "S", 32768
    m_total = 131072; m = m_total // world; tol = 1e-2
else:
    dt, side, uplo, trans, func, n = torch.float64, "L", "L", "N", "M", 32768
    m = 8192; m_total = m * world; tol = 1e-13
g = torch.Generator(device=dev).manual_seed(1234 + 4)
A = torch.empty((n, n), dtype=dt, device=dev).t()
if rank == 0:
    for c0 in range(0, n, 4096):   # built in column slabs to bound the FP32 temporaries
        blk = (2 * torch.rand(n, 4096, dtype=torch.float32, device=dev, generator=g) - 1) / n ** 0.5
        A[:, c0:c0 + 4096] = blk.to(dt)
    A.copy_(torch.tril(A, -1) if uplo == "L" else torch.triu(A, 1))
    A.diagonal().copy_((1 + torch.rand(n, dtype=torch.float32, device=dev, generator=g)).to(dt))
else:
    A.zero_()
gb = torch.Generator(device=dev).manual_seed(777 + rank)
shape = (n, m) if side == "L" else (m, n)
B0 = torch.empty(shape[::-1], dtype=dt, device=dev).t()
B0.copy_((torch.rand(shape, dtype=torch.float32, device=dev, generator=gb) + 1).to(dt))
X = torch.empty(shape[::-1], dtype=dt, device=dev).t()

def step():
    if world > 1 and not a.no_pipeline:
        kw = {} if a.triangle < 0 else {"triangle_only": bool(a.triangle)}
        sh.unified_rectrxm_pipelined(side, uplo, trans, 1.0, func, A, X, src=0, panels=8, handle=h, **kw)
    else:
        if world > 1:
            dist.broadcast(A.t(), src=0)
        nla.unified_rectrxm(side, uplo, trans, 1.0, func, A, X, handle=h)

def sync():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier(); torch.cuda.synchronize()

for _ in range(a.warmup):
    X.copy_(B0); step()
sync()
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
for k in range(a.steps):
    X.copy_(B0); sync()
    ev[k][0].record(); step(); ev[k][1].record()
sync()
tot = torch.tensor([sum(x.elapsed_time(y) for x, y in ev)], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(tot, op=dist.ReduceOp.MAX)
ms = tot.item() / a.steps
# backward error on 256 vectors of this rank, FP64
Ad = (torch.tril(A) if uplo == "L" else torch.triu(A)).double()
if side == "L":
    Xs, Bs = X[:, :256].double(), B0[:, :256].double()
    R = (Ad @ Xs - Bs) if func == "S" else (Xs - Ad @ Bs)
else:
    Xs, Bs = X[:256, :].double(), B0[:256, :].double()
    R = (Xs @ Ad - Bs) if func == "S" else (Xs - Bs @ Ad)
den = (torch.linalg.norm(Ad) * torch.linalg.norm(Xs) + torch.linalg.norm(Bs)) if func == "S" else torch.linalg.norm(Ad) * torch.linalg.norm(Bs)
err = torch.tensor([(torch.linalg.norm(R) / den).item()], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(err, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"config": a.config, "n_gpus": world, "dtype": str(dt), "case": side + uplo + trans + func, "n": n, "rhs_total": m_total, "rhs_per_gpu": m,
                      "ms_per_step": round(ms, 3), "tflops_total": round(float(n) * n * m_total / ms * 1e-9, 1), "broadcast": "blocking" if a.no_pipeline else "8 panels pipelined", "triangle_only": a.triangle,
                      "backward_error_max_over_ranks": err.item(), "tolerance": tol}), flush=True)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
assert err.item() < tol
