import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from neddf_b200 import _lib as L
dev = torch.device("cuda:0")
lib = L.lib()
n, k = int(sys.argv[1]), int(sys.argv[2])
g = torch.Generator().manual_seed(7)
a = torch.randn(256, k, generator=g); b = torch.randn(n, k, generator=g)
ad, bd = a.to(dev), b.to(dev)
c = torch.zeros(256, n, device=dev); cyc = torch.zeros(1, dtype=torch.int64, device=dev)
L.check(lib.neddf_tc_pair_selftest(L.ptr(ad), L.ptr(bd), n, k, L.ptr(c), L.ptr(cyc), 1, L.stream_ptr(dev)))
torch.cuda.synchronize()
ref = a.double() @ b.double().T
cc = c.cpu().double()
print(f"KC={os.environ.get('NEDDF_PAIR_KC')} ROW0={os.environ.get('NEDDF_PAIR_ROW0')} n={n} k={k}: rel err {float((cc - ref).abs().max() / ref.abs().max()):.2e}")
# which B row does each output column look like?
bn = b.double() / b.double().norm(dim=1, keepdim=True)
x = torch.linalg.lstsq(a.double(), cc).solution  # [k, n]: operand each column saw
xn = x / (x.norm(dim=0, keepdim=True) + 1e-30)
m = (bn @ xn).argmax(0)
print("column -> B row:", m.tolist())
