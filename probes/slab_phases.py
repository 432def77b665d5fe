# Probe: in-kernel globaltimer stamps of the fused FP64 slab kernel (option tc_dbg): per 128-row block row, main loop / diagonal solve / write-back.
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as ge
nla = ge.load_package(); h = nla.default_handle(0)
T = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
m = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
func = sys.argv[3] if len(sys.argv) > 3 else "S"
g = torch.Generator(device="cuda").manual_seed(1)
A = (2 * torch.rand(T, T, dtype=torch.float64, device="cuda", generator=g) - 1) / T ** 0.5
A = (torch.tril(A, -1) + torch.diag(1 + torch.rand(T, dtype=torch.float64, device="cuda", generator=g))).t().contiguous().t()
B0 = (torch.rand(T, m, dtype=torch.float64, device="cuda", generator=g) + 1).t().contiguous().t()
X = B0.clone(memory_format=torch.preserve_format)
h.set_option("macro", T); h.set_option("streams", 1)
if len(sys.argv) > 5: h.set_option("slab_kind", int(sys.argv[5]))
dbg = torch.zeros(4096, dtype=torch.int64, device="cuda")
for rep in range(2):
    X.copy_(B0); torch.cuda.synchronize()
    h.set_option("tc_dbg", dbg.data_ptr() if rep == 1 else 0)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); nla.unified_rectrxm("L", "L", "N", 1.0, func, A, X); e1.record(); torch.cuda.synchronize()
    print("kernel ms", e0.elapsed_time(e1))
h.set_option("tc_dbg", 0)
st = dbg.cpu().numpy()
nb = T // 128
t0 = st[0]
tot_main = tot_diag = tot_wb = 0
for r in range(nb):
    a, b, c, d = st[4 * r:4 * r + 4]
    nxt = st[4 * r + 4] if r + 1 < nb else d
    main, diag, wb = b - a, c - b, d - c
    tot_main += main; tot_diag += diag; tot_wb += wb
    W = int(sys.argv[4]) if len(sys.argv) > 4 else 112
    ideal = r * 128 * W * 128 / 64 / 1.965  # ns for r K-blocks of a W-vector CTA at 64 FMA/clk, 1.965 GHz
    print(f"row {r:2d}: main {main/1e3:8.1f} us (ideal {ideal/1e3:7.1f}, eff {ideal/max(main,1):.2f})  diag {diag/1e3:6.1f} us  writeback {wb/1e3:5.1f} us")
print(f"total: main {tot_main/1e3:.1f} us, diag {tot_diag/1e3:.1f} us, writeback {tot_wb/1e3:.1f} us, span {(st[4*nb-1]-t0)/1e3:.1f} us")

