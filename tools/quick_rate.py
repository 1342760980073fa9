"""Quick throughput probe: render_pixels on a slice of the bench frame (tensor-core engine)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import neddf_b200

n_rays = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
engine = sys.argv[3] if len(sys.argv) > 3 else "tc"
dev = torch.device("cuda:0")
sd, _ = bench.seeded_state_dict()
R, T, calib = bench.synthetic_pose(0)
cam = neddf_b200.Camera.from_matrix(neddf_b200.PinholeCalib(calib), R, T).to(dev); cam.update_transform()
first = (bench.H // 2) * bench.W
render = neddf_b200.NeRFRender(network_config=bench.NET_CFG, **bench.RENDER_CFG)
render.load_state_dict(sd); render.to(dev); render.set_iter(-1); render.set_engine(engine)
out = render.render_pixels(bench.W, bench.H, cam, ["color", "depth"], 1, first, n_rays)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    out = render.render_pixels(bench.W, bench.H, cam, ["color", "depth"], 1, first, n_rays)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"{engine}: {ms:.1f} ms per {n_rays} rays -> {n_rays * bench.EVALS_PER_RAY / ms * 1e3:.3e} evaluations/s", flush=True)
