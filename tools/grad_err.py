"""Gradient parity of the training path per engine: parameter gradients of the CUDA path against
autograd through the oracle in fp32 and fp64 (arbiter), worst tensor-normalised error per engine."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import neddf_oracle as orc
from tests.helpers import Case, nerr
import tests.gpu_util as G

for name in ("train", "bunny"):
    c = Case(name)
    n_rays = 6
    d, o = orc.make_rays(c.t("uv")[:n_rays], c.cam)
    dists = c.t("dists_fine")[:n_rays, ::3].contiguous()
    pos, dd, var = orc.make_samples(c.rc, d, o, dists)
    g = torch.Generator().manual_seed(4)
    B, S = dists.shape
    gd, gc, gp = torch.randn(B, S, generator=g), torch.randn(B, S, 3, generator=g), torch.randn(B, S, generator=g)

    def run(dt):
        Pg = {k: v.clone().to(dt).requires_grad_(True) for k, v in c.p_fine.items()}
        ref = orc.field_forward(Pg, c.fc, c.st, pos.to(dt), dd.contiguous().to(dt), var.to(dt))
        ((ref["density"] * gd.to(dt)).sum() + (ref["color"] * gc.to(dt)).sum() + (ref["fields_penalty"] * gp.to(dt)).sum()).backward()
        return {k: v.grad for k, v in Pg.items()}

    g32, g64 = run(torch.float32), run(torch.float64)
    for engine in sys.argv[1:] or ["fp32", "tc", "tc2"]:
        render = G.build_render(c, engine)
        net = render.network_fine
        out = net.forward_rays(d.contiguous().to(G.DEV), o.contiguous().to(G.DEV), dists.to(G.DEV), c.rc.sampling_type,
                               render._ray_radius)
        loss = (out["density"] * gd.to(G.DEV)).sum() + (out["color"] * gc.to(G.DEV)).sum() + (out["fields_penalty"] * gp.to(G.DEV)).sum()
        net.zero_grad()
        loss.backward()
        worst32, worst64, wk = 0.0, 0.0, None
        for k in g32:
            mod, attr = k.rsplit(".", 1)
            obj = net
            for part in mod.split("."):
                obj = obj[int(part)] if part.isdigit() else getattr(obj, part)
            got = getattr(obj, attr).grad.cpu().numpy()
            e32, e64 = nerr(got, g32[k].numpy()), nerr(got, g64[k].numpy())
            if e64 > worst64:
                worst64, wk = e64, k
            worst32 = max(worst32, e32)
        print(f"{name} {engine}: worst gradient error vs fp32 oracle {worst32:.2e}, vs fp64 {worst64:.2e} ({wk})", flush=True)
