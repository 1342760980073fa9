"""Per-step phase durations of the TC megakernel (CTA 0), from the in-kernel clock64() stamps."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import torch
import bench, neddf_b200
from neddf_b200 import _lib as L

dev = torch.device("cuda:0")
sd, _, _ = bench.seeded_state_dict()
render = neddf_b200.NeRFRender(network_config=bench.NET_CFG, **bench.RENDER_CFG)
render.load_state_dict(sd); render.to(dev); render.set_iter(-1); render.set_engine("tc")
net = render.network_fine
R, T, calib = bench.synthetic_pose(0)
cam = neddf_b200.Camera.from_matrix(neddf_b200.PinholeCalib(calib), R, T).to(dev); cam.update_transform()
n_rays = 148 * 32 * 4 // 65 + 1
first = (bench.H // 2) * bench.W
render.render_pixels(bench.W, bench.H, cam, ["color"], 1, first, 4096)  # warm-up, packs weights
buf = torch.zeros(6 * 12 * 8, dtype=torch.int64, device=dev)
h = net._field(dev)
L.check(L.lib().neddf_field_set_timeline(h, L.ptr(buf), buf.numel()))
render.render_pixels(bench.W, bench.H, cam, ["color"], 1, first, 4096)
torch.cuda.synchronize()
L.check(L.lib().neddf_field_set_timeline(h, None, 0))
t = buf.cpu().view(-1, 6)
t = t[t[:, 0] != 0]
names = ["L0", "L1", "L2", "L3", "L4", "L5", "L6", "HDA", "C0", "C1", "C2", "HCOL"]
print("step   mma_issue  wait_full  mma_start->epi_start  epilogue   epi_done->next_mma_start")
for i in range(len(t) - 1):
    a, b, c, d, wf, _ = [int(x) for x in t[i]]
    nxt = int(t[i + 1][0])
    print(f"{names[i % 12]:5s} {b - a:9d} {wf:9d} {c - a:18d} {d - c:12d} {nxt - d:12d}")
tile = int(t[12][0] - t[0][0]) if len(t) > 12 else 0
print("cycles per tile:", tile)
