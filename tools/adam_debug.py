import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from neddf_b200 import losses, optim
from tests.helpers import Case, nerr
import tests.gpu_util as G
c = Case("train")
cam = G.build_camera(c)
uv = c.t("uv").to(G.DEV)
u = (c.t("u_coarse").to(G.DEV), c.t("u_fine").to(G.DEV))
tgt = {"color": torch.rand(uv.shape[0], 3, generator=torch.Generator().manual_seed(1)).to(G.DEV), "mask": torch.ones(uv.shape[0], device=G.DEV)}
loss_fn = losses.RenderLoss()
renders = [G.build_render(c, "auto") for _ in range(3)]
opts = [torch.optim.Adam(renders[0].get_parameters_list(), lr=5e-4), optim.FusedAdam.for_render(renders[1], lr=5e-4),
        torch.optim.Adam(renders[2].get_parameters_list(), lr=5e-4)]
for it in range(3):
    gs = []
    for r, o in zip(renders, opts):
        r.set_iter(c.iter + it)
        out = r.render_rays(uv, cam, uniforms=u)
        loss = torch.sum(torch.stack(list(loss_fn(out, tgt).values())))
        o.zero_grad(set_to_none=True)
        loss.backward()
        gs.append({n: p.grad.clone() for n, p in r.named_parameters()})
        o.step()
    worst = max((nerr(gs[1][n].cpu().numpy(), gs[0][n].cpu().numpy()), n) for n in gs[0])
    worst2 = max((nerr(gs[2][n].cpu().numpy(), gs[0][n].cpu().numpy()), n) for n in gs[0])
    pw = max((nerr(p1.detach().cpu().numpy(), p0.detach().cpu().numpy()), n0) for (n0, p0), (n1, p1) in zip(renders[0].named_parameters(), renders[1].named_parameters()))
    pw2 = max((nerr(p1.detach().cpu().numpy(), p0.detach().cpu().numpy()), n0) for (n0, p0), (n1, p1) in zip(renders[0].named_parameters(), renders[2].named_parameters()))
    print(f"it {it}: loss {float(loss):.6f}; grads fused-vs-torch worst {worst}; torch-vs-torch {worst2}; params fused {pw}; torch twin {pw2}")
