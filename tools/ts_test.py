import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from neddf_b200 import _lib as L
dev = torch.device("cuda:0")
for k in (16, 64, 256):
    g = torch.Generator().manual_seed(k)
    a = torch.randn(128, k, generator=g); b = torch.randn(128, k, generator=g)
    ad, bd = a.to(dev), b.to(dev)
    c = torch.zeros(128, 128, device=dev); cyc = torch.zeros(1, dtype=torch.int64, device=dev)
    for reps in (1, 32):
        L.check(L.lib().neddf_tc_selftest_ts(L.ptr(ad), L.ptr(bd), k, L.ptr(c), L.ptr(cyc), reps, L.stream_ptr(dev)))
        torch.cuda.synchronize()
        ref = a.double() @ b.double().T
        err = float((c.cpu().double() - ref).abs().max() / ref.abs().max())
        n_mma = reps * (k // 16) * 3
        print(f"k={k} reps={reps} rel err {err:.2e}  cycles/MMA {cyc.item() / n_mma:.1f}")
