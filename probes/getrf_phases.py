"""Per-phase clock totals of the cluster panel kernel (CTA 0, thread 0) for one 64-column panel."""
import sys, torch
sys.path.insert(0, ".")
import __graft_entry__ as ge
nla = ge.load_package()
h = nla.default_handle(0)
dbg = torch.zeros(8, dtype=torch.int64, device="cuda")
for m in (64, 1024, 4096):
    A = (torch.rand(64, m, device="cuda", dtype=torch.float64) - 0.5).t()      # m x 64 column-major
    nla.getrf2(A.clone(memory_format=torch.preserve_format)); torch.cuda.synchronize()
    h.set_option("tc_dbg", dbg.data_ptr())
    B = A.clone(memory_format=torch.preserve_format)
    nla.getrf2(B); torch.cuda.synchronize()
    h.set_option("tc_dbg", 0)
    ph = dbg.cpu().numpy()[:6]
    names = ["scan+shuffle", "syncthreads A", "wv reduce + push", "cluster barrier", "cand reduce + swap + sync", "update"]
    print("m", m, "total cycles/column", ph.sum() / 64, {n: round(float(x) / 64) for n, x in zip(names, ph)}, flush=True)
