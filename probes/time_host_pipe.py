# Probe: end-to-end time of the host-buffer paths on ONE GPU (C2 shape): nla_rectrxm_host vs sharded.unified_rectrxm_pipelined_host
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as ge
from importlib import import_module
nla = ge.load_package(); sh = import_module(nla.__name__ + ".sharded"); h = nla.default_handle(0)
n = m = 16384; dt = torch.float64
g = torch.Generator(device="cuda").manual_seed(1)
A = torch.empty((n, n), dtype=dt, device="cuda").t()
A.copy_((2 * torch.rand(n, n, dtype=dt, device="cuda", generator=g) - 1) / n ** 0.5)
A.copy_(torch.tril(A, -1) + torch.diag(1 + torch.rand(n, dtype=dt, device="cuda", generator=g)))
B0 = torch.empty((m, n), dtype=dt, device="cuda").t(); B0.copy_(torch.rand(n, m, dtype=dt, device="cuda", generator=g) + 1)
hostA = torch.empty((n, n), dtype=dt, pin_memory=True); hostA.copy_(A.t())
hostB = torch.empty((m, n), dtype=dt, pin_memory=True); hostB.copy_(B0.t())
hostX = torch.empty((m, n), dtype=dt, pin_memory=True)
dA = torch.empty((n, n), dtype=dt, device="cuda").t(); dX = torch.empty((m, n), dtype=dt, device="cuda").t()
lib = nla.load_library()
def t_c():
    hostX.copy_(hostB); torch.cuda.synchronize(); t0 = time.perf_counter()
    rc = lib.nla_rectrxm_host(h._h, b"L", b"L", b"N", b"S", 0, n, m, 1.0, hostA.data_ptr(), n, hostX.data_ptr(), n); assert rc == 0
    return time.perf_counter() - t0
def t_py(panels):
    hostX.copy_(hostB); torch.cuda.synchronize(); t0 = time.perf_counter()
    sh.unified_rectrxm_pipelined_host("L", "L", "N", 1.0, "S", dA, hostA.t(), hostX.t(), panels=panels)
    torch.cuda.synchronize(); return time.perf_counter() - t0
def t_copy():
    torch.cuda.synchronize(); t0 = time.perf_counter(); dX.t().copy_(hostB, non_blocking=True); torch.cuda.synchronize(); a = time.perf_counter() - t0
    t0 = time.perf_counter(); hostX.copy_(dX.t(), non_blocking=True); torch.cuda.synchronize(); return a, time.perf_counter() - t0
print(json.dumps({"h2d_2GiB_ms": t_copy()[0] * 1e3, "d2h_2GiB_ms": t_copy()[1] * 1e3}))
t_c(); print(json.dumps({"path": "nla_rectrxm_host", "ms": min(t_c() for _ in range(3)) * 1e3}))
for panels in (4, 8, 16):
    t_py(panels); print(json.dumps({"path": "pipelined_host (hostb_gated)", "panels": panels, "ms": min(t_py(panels) for _ in range(3)) * 1e3}), flush=True)
R = torch.tril(A) @ hostX.cuda().t() - B0
print("berr", (torch.linalg.norm(R) / (torch.linalg.norm(torch.tril(A)) * torch.linalg.norm(hostX) + torch.linalg.norm(B0))).item())
