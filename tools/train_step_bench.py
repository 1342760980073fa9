"""Time the training inner loop (BASELINE.json config 4: forward + backward through the field, 1 GPU):
render_rays under autograd on B random pixels, the reference's loss terms, backward, Adam step."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench, neddf_b200

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dev = torch.device("cuda:0")
sd, _ = bench.seeded_state_dict()
render = neddf_b200.NeRFRender(network_config=bench.NET_CFG, **bench.RENDER_CFG)
render.load_state_dict(sd); render.to(dev); render.check_nan = False
R, T, calib = bench.synthetic_pose(0)
cam = neddf_b200.Camera.from_matrix(neddf_b200.PinholeCalib(calib), R, T).to(dev); cam.update_transform()
opt = torch.optim.Adam(render.get_parameters_list(), lr=5e-4)
g = torch.Generator().manual_seed(0)
target = torch.rand(B, 3, generator=g).to(dev)

def step(it):
    render.set_iter(it)
    uv = torch.stack([torch.randint(0, bench.W, (B,), generator=g), torch.randint(0, bench.H, (B,), generator=g)], 1).to(dev)
    out = render.render_rays(uv, cam)
    loss = ((out["color"] - target) ** 2).mean() + 0.1 * ((out["color_coarse"] - target) ** 2).mean() \
        + 0.01 * out["fields_penalty"].mean() + 0.001 * out["fields_penalty_coarse"].mean() \
        + 0.05 * (out["transmittance"] ** 2).mean()
    opt.zero_grad(set_to_none=True)
    loss.backward()
    opt.step()
    return float(loss.detach())

for it in range(3):
    l0 = step(it)
torch.cuda.synchronize()
t0 = time.perf_counter()
K = 5
for it in range(3, 3 + K):
    l1 = step(it)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / K
ev = B * bench.EVALS_PER_RAY
print(f"train step B={B} rays: {dt*1e3:.1f} ms/step, {ev/dt:.3e} evaluations/s (forward+backward+Adam), loss {l0:.4f} -> {l1:.4f}, "
      f"peak memory {torch.cuda.max_memory_allocated()/2**30:.1f} GiB")
