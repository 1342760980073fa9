import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from neddf_b200 import _lib as L
dev = torch.device("cuda:0")
out = torch.zeros(2, dtype=torch.int64, device=dev)
reps = 64
for (a_mn, b_mn, swz, n, label) in [(0, 1, 0, 128, "A K-major / B MN-major, no swizzle, N=128 (megakernel hidden layers)"),
                                    (0, 1, 0, 256, "same, N=256"), (0, 1, 0, 64, "same, N=64"), (0, 1, 0, 32, "same, N=32"),
                                    (0, 0, 0, 128, "A K / B K, no swizzle, N=128"), (1, 1, 0, 128, "A MN / B MN, no swizzle, N=128"),
                                    (1, 0, 0, 16, "A MN / B K, no swizzle, N=16 (heads)"),
                                    (0, 0, 2, 128, "A K / B K, SWIZZLE_128B, N=128"), (0, 0, 2, 256, "A K / B K, SWIZZLE_128B, N=256"),
                                    (0, 1, 2, 128, "A K / B MN, SWIZZLE_128B, N=128"),
                                    (0, 1, 8, 128, "megakernel pattern: N=256 + N=128 per chunk (2 MMAs counted as 2)"),
                                    (0, 1, 9, 128, "pattern with both N=256"),
                                    (0, 1, 10, 128, "N=256 + N=128, commit per chunk"),
                                    (0, 1, 11, 128, "both N=256, commit per chunk")]:
    for _ in range(2):
        L.check(L.lib().neddf_tc_mma_bench(a_mn, b_mn, swz, n, reps, L.ptr(out), L.stream_ptr(dev)))
        torch.cuda.synchronize()
    t = out.cpu().tolist()
    if swz >= 8: t = [x / 2 for x in t]
    print(f"{label:75s} issue {t[0]/(reps*16):7.1f} cyc/MMA   complete {t[1]/(reps*16):7.1f} cyc/MMA   (floor {128*n/256:.0f})")
