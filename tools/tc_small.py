import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import neddf_b200
from oracle import neddf_oracle as orc
from tests.helpers import Case, nerr
import tests.gpu_util as G
c = Case("bunny")
render = G.build_render(c, "tc")
d, o = orc.make_rays(c.t("uv"), c.cam)
pos, dd, var = orc.make_samples(c.rc, d, o, c.t("dists_fine"))
with torch.no_grad():
    out = render.network_fine(neddf_b200.Sampling(pos.to(G.DEV), dd.contiguous().to(G.DEV), var.to(G.DEV)))
torch.cuda.synchronize()
for k in ("distance", "density", "color", "fields_penalty"):
    print(k, nerr(out[k].cpu().numpy(), c.z["field_fine_" + k]))
