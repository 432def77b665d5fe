import sys, torch
sys.path.insert(0, ".")
import __graft_entry__ as ge
nla = ge.load_package()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
dt = torch.float64 if (len(sys.argv) < 3 or sys.argv[2] == "f64") else torch.float32
g = torch.Generator(device="cuda").manual_seed(n)
A = (torch.rand(n, n, device="cuda", dtype=dt, generator=g) - 0.5).t()
nla.getrf2(A)
torch.cuda.synchronize()
