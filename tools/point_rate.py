"""Point-query fast path: NeDDF.forward(Sampling) on explicit positions, variance 0 - what
`voxelize` (base_neuralfield.py:49-79) and `render_field_slice` (nerf_render.py:263-336) call.  Full output set
(distance, density, colour, penalties, auxiliary gradient => F_full per evaluation) and a 256^3 `voxelize("density")`."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import neddf_b200

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 22
dev = torch.device("cuda:0")
sd, _ = bench.seeded_state_dict()
render = neddf_b200.NeRFRender(network_config=bench.NET_CFG, **bench.RENDER_CFG)
render.load_state_dict(sd); render.to(dev); render.set_iter(-1)
net = render.get_network()
g = torch.Generator().manual_seed(0)
pos = (torch.rand(1, n, 3, generator=g) * 2.2 - 1.1).to(dev)
sdir = torch.zeros_like(pos); sdir[..., 0] = 1.0
s = neddf_b200.Sampling(pos, sdir, torch.zeros_like(pos))
for engine in ("tc", "tc2", "fp32"):
    render.set_engine(engine)
    reps = 1 if engine == "fp32" else 3
    with torch.no_grad():
        net(s); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            out = net(s)
        e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"point query {engine}: {n / ms * 1e3:.3e} evaluations/s (full outputs, {n * bench.F_FULL / ms * 1e-9:.0f} TFLOP/s algorithmic), "
          f"{ms:.1f} ms for {n} points", flush=True)
render.set_engine("auto")
t0 = time.time()
vox = net.voxelize("density", cube_range=1.1, cube_resolution=256, chunk=1 << 21)
print(f"voxelize 256^3 (auto engine, host numpy result): {time.time() - t0:.2f} s, {256 ** 3 / (time.time() - t0):.3e} points/s end to end")
