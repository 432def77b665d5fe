# Probe: nla_rectrxm_host (C2 shape, pinned host buffers): fused-slab block orders at the ends / in the middle of the diagonal, RHS slabs.
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as ge
nla = ge.load_package(); h = nla.default_handle(0); lib = nla.load_library()
n = m = 16384; dt = torch.float64
g = torch.Generator(device="cuda").manual_seed(1)
A = torch.empty((n, n), dtype=dt, device="cuda").t()
A.copy_((2 * torch.rand(n, n, dtype=dt, device="cuda", generator=g) - 1) / n ** 0.5)
A.copy_(torch.tril(A, -1) + torch.diag(1 + torch.rand(n, dtype=dt, device="cuda", generator=g)))
hostA = torch.empty((n, n), dtype=dt, pin_memory=True); hostA.copy_(A.t())
hostB = torch.empty((m, n), dtype=dt, pin_memory=True); hostB.uniform_(1, 2)
hostX = torch.empty((m, n), dtype=dt, pin_memory=True)
def run():
    hostX.copy_(hostB); torch.cuda.synchronize(); t0 = time.perf_counter()
    rc = lib.nla_rectrxm_host(h._h, b"L", b"L", b"N", b"S", 0, n, m, 1.0, hostA.data_ptr(), n, hostX.data_ptr(), n); assert rc == 0
    return (time.perf_counter() - t0) * 1e3
for edge, mid, slabs in ((1024, 1024, 0), (1024, 2048, 0), (1024, 4096, 0), (512, 4096, 0), (1024, 4096, 8), (512, 2048, 0), (256, 4096, 0)):
    h.set_option("host_macro", edge); h.set_option("host_macro_mid", mid); h.set_option("host_slabs", slabs)
    run(); ts = [run() for _ in range(4)]
    print(json.dumps({"host_macro": edge, "host_macro_mid": mid, "host_slabs": slabs, "wall_ms_min": round(min(ts), 2), "all": [round(t, 1) for t in ts]}), flush=True)
R = torch.tril(A) @ hostX.cuda().t() - hostB.cuda().t()
print("berr", (torch.linalg.norm(R) / (torch.linalg.norm(torch.tril(A)) * torch.linalg.norm(hostX) + torch.linalg.norm(hostB))).item())
