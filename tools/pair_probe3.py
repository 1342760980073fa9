import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from neddf_b200 import _lib as L
dev = torch.device("cuda:0")
lib = L.lib()
n, k = 128, 64
g = torch.Generator().manual_seed(7)
a = torch.randn(256, k, generator=g); b = torch.randn(n, k, generator=g)
ad, bd = a.to(dev), b.to(dev)
ref = a.double() @ b.double().T
os.environ["NEDDF_PAIR_KC"] = "96"
for boff in (0, 65536, 98304, 131072):
    os.environ["NEDDF_PAIR_BASE"] = str(boff)
    mode = boff
    c = torch.zeros(256, n, device=dev); cyc = torch.zeros(1, dtype=torch.int64, device=dev)
    L.check(lib.neddf_tc_pair_selftest(L.ptr(ad), L.ptr(bd), n, k, L.ptr(c), L.ptr(cyc), 1, L.stream_ptr(dev)))
    torch.cuda.synchronize()
    cc = c.cpu().double()
    err = float((cc - ref).abs().max() / ref.abs().max())
    al = []
    for col in range(32):
        m = [c2 for c2 in range(n) if float((cc[:, col] - ref[:, c2]).abs().max()) < 1e-3]
        al.append(m[0] if m else -1)
    print(f"B buffers at +{mode}: rel err {err:.2e}; output column -> true column (first 32): {al}", flush=True)
