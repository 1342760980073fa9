"""Per-chunk stamps of the MMA warp of the pair kernel (NEDDF_TC2_DEBUG & 512), second tile of cluster 0:
time spent per chunk-pass (wait for the chunk + issue)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench, neddf_b200
from neddf_b200 import _lib as L
dev = torch.device("cuda:0")
sd, _ = bench.seeded_state_dict()
render = neddf_b200.NeRFRender(network_config=bench.NET_CFG, **bench.RENDER_CFG)
render.load_state_dict(sd); render.to(dev); render.set_iter(-1); render.set_engine("tc2"); render.check_nan = False
net = render.network_fine
R, T, calib = bench.synthetic_pose(0)
cam = neddf_b200.Camera.from_matrix(neddf_b200.PinholeCalib(calib), R, T).to(dev); cam.update_transform()
first = (bench.H // 2) * bench.W
render.render_pixels(bench.W, bench.H, cam, ["color"], 1, first, 65536)
h = net._field(dev)
os.environ["NEDDF_TC2_DEBUG"] = "512"
buf = torch.zeros(4 * 400, dtype=torch.int64, device=dev)
L.check(L.lib().neddf_field_set_timeline(h, L.ptr(buf), buf.numel()))
render.render_pixels(bench.W, bench.H, cam, ["color"], 1, first, 65536)
torch.cuda.synchronize()
L.check(L.lib().neddf_field_set_timeline(h, None, 0))
t = buf.cpu().view(-1, 4)
t = t[t[:, 0] != 0]
names = ["L0", "L1", "L2", "L3", "L4", "L5", "L6", "C0", "C1", "C2"]
print("step hs stage  wait+issue  gap_since_prev_chunk_end")
for i in range(len(t)):
    a, b, tag, _ = [int(x) for x in t[i]]
    si, hs, stage = tag // 16, (tag // 8) & 1, tag & 7
    prev = int(t[i - 1][1]) if i else a
    print(f"{names[si]:3s} {hs} {stage}  {b - a:8d} {a - prev:8d}")
