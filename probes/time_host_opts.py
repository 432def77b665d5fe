# Probe: nla_rectrxm_host (C2 shape, pinned host buffers) under different fused-slab cutoffs; per-op device time vs wall time.
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
for macro, slabs in ((2048, 0), (2048, 1), (2048, 2), (1024, 0)):
    h.set_option("macro", macro); h.set_option("host_slabs", slabs)
    run(); ms = min(run() for _ in range(3))
    h.set_option("profile", 1); run(); prof = h.profile_read(); h.set_option("profile", 0)
    gm = sum(x[2] for x in prof if x[0] == 1); lf = sum(x[2] for x in prof if x[0] == 0); span = sum(x[2] for x in prof if x[0] == 2)
    print(json.dumps({"macro": macro, "host_slabs": slabs, "wall_ms": round(ms, 2), "ops": len(prof), "first_to_last_launch_ms": round(span, 2), "gemm_ms": round(gm, 2), "leaf_ms": round(lf, 2), "device_ms_sum": round(gm + lf, 2)}), flush=True)
h.set_option("macro", 2048); h.set_option("host_slabs", 0); run()
R = torch.tril(A) @ hostX.cuda().t() - hostB.cuda().t()
print("berr", (torch.linalg.norm(R) / (torch.linalg.norm(torch.tril(A)) * torch.linalg.norm(hostX) + torch.linalg.norm(hostB))).item())
