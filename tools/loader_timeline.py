"""Loader-warp stamps of the pair kernel (NEDDF_TC2_DEBUG & 128): per own chunk of loader warp 17 of one CTA:
wait for the ring stage, tcgen05.st + wait, arrive."""
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
render.render_pixels(bench.W, bench.H, cam, ["color"], 1, first, 4096)
h = net._field(dev)
extra = int(sys.argv[1]) if len(sys.argv) > 1 else 0
for blk in (0, 1):
    os.environ["NEDDF_TC2_DEBUG"] = str(128 + extra + (blk << 8))
    buf = torch.zeros(4 * 400, dtype=torch.int64, device=dev)
    L.check(L.lib().neddf_field_set_timeline(h, L.ptr(buf), buf.numel()))
    render.render_pixels(bench.W, bench.H, cam, ["color"], 1, first, 8192)
    torch.cuda.synchronize()
    L.check(L.lib().neddf_field_set_timeline(h, None, 0))
    t = buf.cpu().view(-1, 4)
    t = t[t[:, 0] != 0]
    print(f"== CTA {blk} (rank {blk}), loader warp 17: chunk  wait_empty  st+wait  arrive  period")
    for i in range(160, min(len(t), 200)):
        a, b, c, d = [int(x) for x in t[i]]
        print(f"{2 * i:5d} {b - a:9d} {c - b:8d} {d - c:6d} {a - int(t[i - 1][0]):8d}")
