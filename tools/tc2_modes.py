"""Stage-isolation timing of the pair kernel: whole-kernel rate with stages switched off
(NEDDF_TC2_DEBUG bits; results are garbage for modes >= 4)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench, neddf_b200
modes = [int(a) for a in sys.argv[1:]] or [0, 4, 8, 16, 32, 20, 28, 64, 80]
dev = torch.device("cuda:0")
sd, _ = bench.seeded_state_dict()
R, T, calib = bench.synthetic_pose(0)
cam = neddf_b200.Camera.from_matrix(neddf_b200.PinholeCalib(calib), R, T).to(dev); cam.update_transform()
first = (bench.H // 2) * bench.W
render = neddf_b200.NeRFRender(network_config=bench.NET_CFG, **bench.RENDER_CFG)
render.load_state_dict(sd); render.to(dev); render.set_iter(-1); render.set_engine(os.environ.get("ENGINE", "tc2")); render.check_nan = False
n_rays = 65536
render.render_pixels(bench.W, bench.H, cam, ["color", "depth"], 1, first, n_rays)
groups = [int(g) for g in os.environ.get("TC2_GROUPS", "8").split(",")]
for m in [(mm, g) for g in groups for mm in modes]:
    os.environ["NEDDF_TC2_GROUP"] = str(m[1])
    m = m[0]
    os.environ["NEDDF_TC2_DEBUG"] = str(m)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(2):
        render.render_pixels(bench.W, bench.H, cam, ["color", "depth"], 1, first, n_rays)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 2
    print(f"group {os.environ['NEDDF_TC2_GROUP']} debug {m:3d}: {n_rays * bench.EVALS_PER_RAY / ms * 1e3:.3e} evaluations/s", flush=True)
