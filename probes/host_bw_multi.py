# Probe (torchrun, N ranks): concurrent host<->device bandwidth per rank (contiguous 2 GiB copies) and the B-streaming host pipeline
# (nla_rectrxm_hostb_gated with A already resident) run by all ranks at once.  Explains the end-to-end scaling of bench.py.
import os, sys, time, json, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import __graft_entry__ as ge
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1: dist.init_process_group("nccl", device_id=dev)
nla = ge.load_package(); h = nla.Handle(local); lib = nla.load_library()
n = m = 16384; dt = torch.float64
def barrier():
    torch.cuda.synchronize()
    if world > 1: dist.barrier(); torch.cuda.synchronize()
def tmax(v):
    t = torch.tensor([v], dtype=torch.float64, device=dev)
    if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()
hB = torch.empty((m, n), dtype=dt, pin_memory=True); hB.uniform_(1, 2)
hX = torch.empty((m, n), dtype=dt, pin_memory=True)
dX = torch.empty((m, n), dtype=dt, device=dev)
out = {"ranks": world}
for name, fn in (("h2d", lambda: dX.copy_(hB, non_blocking=True)), ("d2h", lambda: hX.copy_(dX, non_blocking=True))):
    fn(); barrier(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); out[name + "_2GiB_ms_max"] = round(tmax(time.perf_counter() - t0) * 1e3, 1)
s2 = torch.cuda.Stream()
barrier(); t0 = time.perf_counter()
dX.copy_(hB, non_blocking=True)
with torch.cuda.stream(s2): hX.copy_(dX, non_blocking=True)
torch.cuda.synchronize(); out["h2d+d2h_concurrent_ms_max"] = round(tmax(time.perf_counter() - t0) * 1e3, 1)
# B-streaming pipeline with A resident
g = torch.Generator(device=dev).manual_seed(1)
A = torch.empty((n, n), dtype=dt, device=dev).t()
A.copy_((2 * torch.rand(n, n, dtype=dt, device=dev, generator=g) - 1) / n ** 0.5)
A.copy_(torch.tril(A, -1) + torch.diag(1 + torch.rand(n, dtype=dt, device=dev, generator=g)))
def pipe():
    rc = lib.nla_rectrxm_hostb_gated(h._h, b"L", b"L", b"N", b"S", 0, n, m, 1.0, A.data_ptr(), n, hX.data_ptr(), n, 0, 0, None); assert rc == 0
hX.copy_(hB); pipe()
for rep in range(2):
    hX.copy_(hB); barrier(); t0 = time.perf_counter(); pipe(); torch.cuda.synchronize(); out["hostb_pipeline_ms_max_%d" % rep] = round(tmax(time.perf_counter() - t0) * 1e3, 1)
if rank == 0: print(json.dumps(out), flush=True)
if world > 1: dist.barrier(); dist.destroy_process_group()
