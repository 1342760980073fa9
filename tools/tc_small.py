import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import neddf_b200
from oracle import neddf_oracle as orc
from tests.helpers import Case, nerr
import tests.gpu_util as G
c = Case("bunny")
engine = sys.argv[1] if len(sys.argv) > 1 else "tc"
render = G.build_render(c, engine)
d, o = orc.make_rays(c.t("uv"), c.cam)
pos, dd, var = orc.make_samples(c.rc, d, o, c.t("dists_fine"))
with torch.no_grad():
    out = render.network_fine(neddf_b200.Sampling(pos.to(G.DEV), dd.contiguous().to(G.DEV), var.to(G.DEV)))
torch.cuda.synchronize()
for k in ("distance", "density", "color", "fields_penalty"):
    print(k, nerr(out[k].cpu().numpy(), c.z["field_fine_" + k]))
render.network_fine.check_engine_status()
# images-only path (value rows only in the colour trunk) through render_rays-style forward_rays
uv, u_c = c.t("uv"), c.t("u_coarse")
dists = orc.coarse_dists(c.rc, u_c)
with torch.no_grad():
    o2 = render.network_fine.forward_rays(d.to(G.DEV), o.to(G.DEV), dists.to(G.DEV), c.rc.sampling_type,
                                          neddf_b200.ray.CONE_RAY_RADIUS if c.rc.sampling_type == "cone" else 0.0,
                                          need_penalty=False, need_aux=False)
torch.cuda.synchronize()
pos2, dd2, var2 = orc.make_samples(c.rc, d, o, dists)
with torch.no_grad():
    ref = orc.field_forward(c.p_fine, c.fc, c.st, pos2, dd2, var2)
for k in ("density", "color"):
    print("eval", k, nerr(o2[k].cpu().numpy(), ref[k].numpy()))
