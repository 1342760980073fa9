import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from neddf_b200 import _lib as L
dev = torch.device("cuda:0")
for lbo, sbo in ((128, 256), (256, 128), (16, 256), (128, 4096)):
    out = torch.zeros(128, 8, dtype=torch.int32, device=dev)
    L.check(L.lib().neddf_tc_cp_probe(lbo, sbo, L.ptr(out), L.stream_ptr(dev)))
    torch.cuda.synchronize()
    o = out.cpu().numpy().astype("uint32")
    lo, hi = o & 0xffff, o >> 16
    print(f"lbo={lbo} sbo={sbo}")
    for lane in (0, 1, 7, 8, 9, 64):
        print(f"  lane {lane:3d}: half indices (byte offset/2) per column:", [(int(a), int(b)) for a, b in zip(lo[lane], hi[lane])])
