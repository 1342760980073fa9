"""Tiny image-mode render (the batched colour trunk of the tc engine, TMA copy-back included) for
compute-sanitizer runs: 600 rays of the bench frame through render_pixels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench, neddf_b200
engine = sys.argv[1] if len(sys.argv) > 1 else "tc"
n_rays = int(sys.argv[2]) if len(sys.argv) > 2 else 600
dev = torch.device("cuda:0")
sd, _ = bench.seeded_state_dict()
render = neddf_b200.NeRFRender(network_config=bench.NET_CFG, **bench.RENDER_CFG)
render.load_state_dict(sd); render.to(dev); render.set_iter(-1); render.set_engine(engine)
R, T, calib = bench.synthetic_pose(0)
cam = neddf_b200.Camera.from_matrix(neddf_b200.PinholeCalib(calib), R, T).to(dev); cam.update_transform()
first = (bench.H // 2) * bench.W
out = render.render_pixels(bench.W, bench.H, cam, ["color", "depth"], 1, first, n_rays)
torch.cuda.synchronize()
print(engine, "image path finite:", bool(torch.isfinite(out["color"]).all()), float(out["color"].mean()))
