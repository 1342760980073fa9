"""Per-step phase durations of the TC megakernel (CTA 0), from the in-kernel clock64() stamps.

usage: timeline.py [debug modes ...]   (NEDDF_TC_DEBUG bit masks, default 0 = the real kernel;
1 = no MMAs, 2 = loaders skip the L2 reads, 4 = loaders skip the TMEM stores, 8 = epilogue skips the
hidden-layer math: they isolate which stage bounds a layer, results are garbage)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import torch
import bench, neddf_b200
from neddf_b200 import _lib as L

modes = [int(a) for a in sys.argv[1:]] or [0]
dev = torch.device("cuda:0")
sd, _ = bench.seeded_state_dict()
render = neddf_b200.NeRFRender(network_config=bench.NET_CFG, **bench.RENDER_CFG)
render.load_state_dict(sd); render.to(dev); render.set_iter(-1); render.set_engine("tc")
render.check_nan = False
net = render.network_fine
R, T, calib = bench.synthetic_pose(0)
cam = neddf_b200.Camera.from_matrix(neddf_b200.PinholeCalib(calib), R, T).to(dev); cam.update_transform()
first = (bench.H // 2) * bench.W
render.render_pixels(bench.W, bench.H, cam, ["color"], 1, first, 4096)  # warm-up, packs weights
h = net._field(dev)
names = ["L0", "L1", "L2", "L3", "L4", "L5", "L6", "HDA", "C0", "C1", "C2", "HCOL"]
for mode in modes:
    os.environ["NEDDF_TC_DEBUG"] = str(mode)
    buf = torch.zeros(6 * 160, dtype=torch.int64, device=dev)
    L.check(L.lib().neddf_field_set_timeline(h, L.ptr(buf), buf.numel()))
    render.render_pixels(bench.W, bench.H, cam, ["color"], 1, first, 4096)
    torch.cuda.synchronize()
    L.check(L.lib().neddf_field_set_timeline(h, None, 0))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    render.render_pixels(bench.W, bench.H, cam, ["color"], 1, first, 32768)
    e1.record(); torch.cuda.synchronize()
    rate = 32768 * bench.EVALS_PER_RAY / e0.elapsed_time(e1) * 1e3
    t = buf.cpu().view(-1, 6)
    t = t[t[:, 0] != 0]
    print(f"== debug mode {mode}: {rate:.3e} evaluations/s")
    print("step   mma_issue  mma_start->epi_start  epilogue   epi_done->next_mma_start")
    # the stamps carry the step index (slot 4: MMA warp, slot 5: epilogue); the program is four trunk passes
    # (L0..HDA) and one colour pass per group when the colour trunk is batched, else all twelve steps per tile
    batch = int(os.environ.get("NEDDF_TC_BATCH", "4"))
    per_group = 8 * batch + 4 if batch > 1 else 12
    lo, hi = per_group, min(len(t) - 1, 2 * per_group)  # second group of the CTA: steady state
    for i in range(lo, hi):
        a, b, c, d, si, _ = [int(x) for x in t[i]]
        nxt = int(t[i + 1][0])
        print(f"{names[si]:5s} {b - a:9d} {c - a:18d} {d - c:12d} {nxt - d:12d}")
    if len(t) > 2 * per_group:
        grp = int(t[2 * per_group][0] - t[per_group][0])
        print(f"cycles per group of {batch if batch > 1 else 1} tile(s): {grp}  = {grp // (batch if batch > 1 else 1)} per tile", flush=True)
