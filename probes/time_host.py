# Probe: times nla_rectrxm_host (pinned host buffers) for option sweeps and raw PCIe copies.
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as ge
nla = ge.load_package(); h = nla.Handle(0); lib = nla.load_library()
n = m = 16384
g = torch.Generator().manual_seed(1)
hA = torch.empty((n, n), dtype=torch.float64, pin_memory=True); hB = torch.empty((m, n), dtype=torch.float64, pin_memory=True); hX = torch.empty((m, n), dtype=torch.float64, pin_memory=True)
A = (2 * torch.rand(n, n, dtype=torch.float64, device="cuda") - 1) / n ** 0.5
A = torch.tril(A, -1) + torch.diag(1 + torch.rand(n, dtype=torch.float64, device="cuda"))
hA.copy_(A.t()); hB.copy_(torch.rand(m, n, dtype=torch.float64, device="cuda") + 1); del A
d = torch.empty((m, n), dtype=torch.float64, device="cuda")
for name, fn in (("h2d", lambda: d.copy_(hB, non_blocking=True)), ("d2h", lambda: hX.copy_(d, non_blocking=True))):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(json.dumps({"copy": name, "GBps": 2 ** 31 / dt * 1e-9}))
for slabs in (0,):
    for macro in (1024, 2048, 4096):
        h.set_option("macro", macro)
        ts = []
        for r in range(3):
            hX.copy_(hB); torch.cuda.synchronize(); t0 = time.perf_counter()
            rc = lib.nla_rectrxm_host(h._h, b"L", b"L", b"N", b"S", 0, n, m, 1.0, hA.data_ptr(), n, hX.data_ptr(), n); assert rc == 0
            ts.append(time.perf_counter() - t0)
        print(json.dumps({"host_slabs": slabs, "macro": macro, "ms": round(min(ts[1:]) * 1e3, 1), "tflops": round(n * n * m / min(ts[1:]) * 1e-12, 2)}), flush=True)
