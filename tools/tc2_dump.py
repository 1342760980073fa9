"""Debug dump of the pair kernel (cluster 0, first tile, step given): AUX operand rows and accumulators."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np
import torch
import neddf_b200
from neddf_b200 import _lib as L
from oracle import neddf_oracle as orc
from tests.helpers import Case
import tests.gpu_util as G

step = int(sys.argv[1]) if len(sys.argv) > 1 else 0
n = 64
c = Case("bunny")
render = G.build_render(c, "tc2")
net = render.network_fine
d, o = orc.make_rays(c.t("uv"), c.cam)
pos, dd, var = orc.make_samples(c.rc, d, o, c.t("dists_fine"))
pos = pos.reshape(-1, 3)[:n].contiguous().to(G.DEV)
dd = dd.reshape(-1, 3)[:n].contiguous().to(G.DEV)
var = var.reshape(-1, 3)[:n].contiguous().to(G.DEV)
h = net._field(G.DEV)
st = net._state_struct()
dump = torch.zeros(2, 6144 + 2 * 128 * 128, device=G.DEV)
L.check(L.lib().neddf_field_set_debug_dump(h, L.ptr(dump), step))
outs = [torch.zeros(n, device=G.DEV) for _ in range(2)] + [torch.zeros(n, 3, device=G.DEV)] + [torch.zeros(n, device=G.DEV) for _ in range(2)]
L.check(L.lib().neddf_field_forward(h, C.byref(st), L.ptr(pos), L.ptr(dd), L.ptr(var), n, L.ptr(outs[0]), L.ptr(outs[1]),
        L.ptr(outs[2]), L.ptr(outs[3]), L.ptr(outs[4]), L.OUT_FULL, L.ENGINE_IDS["tc2"], L.stream_ptr(G.DEV)))
torch.cuda.synchronize()
L.check(L.lib().neddf_field_set_debug_dump(h, None, 0))
dm = dump.cpu()
torch.save(dm, "gpurun_out/tc2_dump.pt")
np.set_printoptions(precision=4, suppress=True, linewidth=220)
for rank in range(2):
    words = dm[rank, :6144].numpy().view(np.uint32)
    halves = words.view(np.float16).astype(np.float32)          # 12288 halves: [row/8][k 96][row%8]
    aux = halves.reshape(16, 96, 8).transpose(0, 2, 1).reshape(128, 96)   # [row][k]
    acc = dm[rank, 6144:].reshape(2, 128, 128).numpy()                    # [hs][lane][col]
    print(f"--- CTA {rank}")
    for row in (0, 1, 16, 32, 48, 64, 80):
        print(f"AUX row {row:3d} k0..11:", aux[row, :12])
    print("acc hs0 lane0 cols 0..3  :", acc[0, 0, :4], " cols 16..19:", acc[0, 0, 16:20], " cols 64..67:", acc[0, 0, 64:68], " cols 80..83", acc[0, 0, 80:84])
    print("acc hs1 lane0 cols 0..3  :", acc[1, 0, :4], " cols 16..19:", acc[1, 0, 16:20])
