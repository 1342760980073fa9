"""Weight-pipeline stamps of the pair kernel (NEDDF_TC2_DEBUG & 128), leader CTA of cluster 0, per chunk:
when the cp issuer started waiting, when the producer had issued the TMA, when it landed, when the peer's
landed, when the tensor-memory stage was free."""
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
os.environ["NEDDF_TC2_DEBUG"] = "128"
buf = torch.zeros(6 * 400, dtype=torch.int64, device=dev)
L.check(L.lib().neddf_field_set_timeline(h, L.ptr(buf), buf.numel()))
render.render_pixels(bench.W, bench.H, cam, ["color"], 1, first, 65536)
torch.cuda.synchronize()
L.check(L.lib().neddf_field_set_timeline(h, None, 0))
t = buf.cpu().view(-1, 6)
t = t[t[:, 0] != 0]
print("chunk  wait_start->landed  tma_latency(issue->landed)  peer_extra  stage_free_extra  period")
for i in range(158, min(len(t), 158 + 79)):
    a, ti, b, c, d, e = [int(x) for x in t[i]]
    print(f"{i:5d} {b - a:12d} {b - ti:18d} {c - b:14d} {d - c:14d} {a - int(t[i - 1][0]):10d}"
          f"   cp x4 {e & 0xfffff:5d}  commit a_full {(e >> 20) & 0xfffff:5d}  commit s_empty {e >> 40:5d}")
