"""Per-step phase durations of the CTA-pair TC megakernel (cluster 0, leader CTA), from the in-kernel
clock64() stamps: MMA issue window of the step, epilogue windows of the two sample halves."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench, neddf_b200
from neddf_b200 import _lib as L

n_rays = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
dev = torch.device("cuda:0")
sd, _ = bench.seeded_state_dict()
render = neddf_b200.NeRFRender(network_config=bench.NET_CFG, **bench.RENDER_CFG)
render.load_state_dict(sd); render.to(dev); render.set_iter(-1); render.set_engine("tc2")
render.check_nan = False
net = render.network_fine
R, T, calib = bench.synthetic_pose(0)
cam = neddf_b200.Camera.from_matrix(neddf_b200.PinholeCalib(calib), R, T).to(dev); cam.update_transform()
first = (bench.H // 2) * bench.W
render.render_pixels(bench.W, bench.H, cam, ["color"], 1, first, 4096)  # warm-up, packs weights
h = net._field(dev)
names = ["L0", "L1", "L2", "L3", "L4", "L5", "L6", "C0", "C1", "C2"]
ns = len(names)
buf = torch.zeros(6 * ns * 8, dtype=torch.int64, device=dev)
L.check(L.lib().neddf_field_set_timeline(h, L.ptr(buf), buf.numel()))
render.render_pixels(bench.W, bench.H, cam, ["color"], 1, first, 8192)
torch.cuda.synchronize()
L.check(L.lib().neddf_field_set_timeline(h, None, 0))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
render.render_pixels(bench.W, bench.H, cam, ["color"], 1, first, n_rays)
e1.record(); torch.cuda.synchronize()
rate = n_rays * bench.EVALS_PER_RAY / e0.elapsed_time(e1) * 1e3
t = buf.cpu().view(-1, 6)
t = t[t[:, 0] != 0]
print(f"== tc2: {rate:.3e} evaluations/s")
print("step   mma_window  start->epi0  epi0   epi0_end->epi1  epi1   epi1_end->next_mma_start")
for i in range(ns, min(len(t) - 1, 3 * ns)):
    a, b, c0, d0, c1, d1 = [int(x) for x in t[i]]
    nxt = int(t[i + 1][0])
    print(f"{names[i % ns]:5s} {b - a:9d} {c0 - a:11d} {d0 - c0:7d} {c1 - d0:12d} {d1 - c1:8d} {nxt - d1:12d}")
if len(t) > 2 * ns:
    print("cycles per pair tile:", int(t[2 * ns][0] - t[ns][0]), flush=True)
