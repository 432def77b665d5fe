# Probe: bitwise run-to-run determinism and exact linearity of the tensor-core path; prints where mismatches sit.
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as ge
nla = ge.load_package(); h = nla.default_handle(0)
dname, case, n, m = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
for kv in sys.argv[5:]:
    k, v = kv.split("="); h.set_option(k, int(v))
dt = getattr(torch, dname); side, uplo, trans, func = case
g = torch.Generator(device="cuda").manual_seed(4321)
A = (2 * torch.rand(n, n, dtype=torch.float32, device="cuda", generator=g) - 1) / n ** 0.5
A = (torch.tril(A, -1) if uplo == "L" else torch.triu(A, 1)) + torch.diag(1 + torch.rand(n, dtype=torch.float32, device="cuda", generator=g))
dA = A.to(dt).t().contiguous().t()
shape = (n, m) if side == "L" else (m, n)
B0 = (torch.rand(shape, dtype=torch.float32, device="cuda", generator=g) + 1).to(dt).t().contiguous().t()
h.set_option("streams", 1)
def run(B):
    X = B.clone(memory_format=torch.preserve_format)
    nla.unified_rectrxm(side, uplo, trans, 1.0, func, dA, X); torch.cuda.synchronize()
    return X
X1 = run(B0); X1b = run(B0); X2 = run((2 * B0).t().contiguous().t())
def where(D):
    idx = D.nonzero()
    if idx.numel() == 0: return {"count": 0}
    r, c = idx[:, 0], idx[:, 1]
    return {"count": int(idx.shape[0]), "rows": [int(r.min()), int(r.max())], "cols": [int(c.min()), int(c.max())],
            "row_mod128_hist": torch.bincount(r % 128, minlength=128)[:8].tolist(), "first": idx[:5].tolist()}
print(json.dumps({"case": sys.argv[1:5], "rerun_mismatch": where(X1 != X1b), "linearity_mismatch": where(X2 != 2 * X1),
                  "lin_maxabs": float((X2.float() - 2 * X1.float()).abs().max()), "min_abs_X": float(X1.float().abs().min())}))
