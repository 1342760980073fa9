"""NeRF field variant: the oracle restatement (oracle.nerf_forward, nerf.py:107-165) against
goldens recorded from the REAL reference (tests/golden/make_nerf_golden.py: the reference's NeRF inside its
NeRFRender, recorded uniforms), per-sample field outputs and the composited render."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import neddf_oracle as orc
from tests.helpers import GOLDEN, nerr


class NerfCase:
    def __init__(self, name: str):
        z = np.load(os.path.join(GOLDEN, f"case_nerf_{name}.npz"), allow_pickle=False)
        self.z = {k: z[k] for k in z.files}
        meta = json.loads(str(self.z["cfg"]))
        self.net_cfg, self.render_cfg, self.iter = meta["net"], meta["render"], int(meta["iter"])
        self.nc = orc.NerfConfig.from_dict(self.net_cfg)
        self.rc = orc.RenderConfig.from_dict(self.render_cfg)
        self.alpha = self.nc.lowpass_alpha_at(self.iter)
        cal = [float(v) for v in self.z["cam_calib"]]
        self.cam = orc.CameraPose(torch.from_numpy(self.z["cam_R"]), torch.from_numpy(self.z["cam_T"]), *cal)

    def params(self, tag: str):
        """[in,out] weights under the reference's state_dict names (torch Linear stores [out,in])."""
        pre = f"w_{tag}." if f"w_{tag}.layers.0.weight" in self.z else "w_fine."
        out = {}
        for k, v in self.z.items():
            if k.startswith(pre):
                t = torch.from_numpy(v)
                out[k[len(pre):]] = t.t().contiguous() if k.endswith(".weight") else t
        return out

    def t(self, k):
        return torch.from_numpy(self.z[k])


@pytest.mark.parametrize("name", ["relu", "tanhexp"])
def test_nerf_oracle_matches_reference(name):
    c = NerfCase(name)
    shapes = {n: (i, o) for n, i, o in orc.nerf_layer_shapes(c.nc)}
    pf = c.params("fine")
    assert {k[:-7] for k in pf if k.endswith(".weight")} == set(shapes)
    assert all(tuple(pf[n + ".weight"].shape) == shapes[n] for n in shapes)
    d, o = orc.make_rays(c.t("uv"), c.cam)
    for tag, dists in (("coarse", orc.coarse_dists(c.rc, c.t("u_coarse"))), ("fine", c.t("dists_fine"))):
        pos, dd, var = orc.make_samples(c.rc, d, o, dists)
        with torch.no_grad():
            out = orc.nerf_forward(c.params(tag), c.nc, c.alpha, pos, dd, var)
        for k in ("density", "color"):
            assert nerr(out[k].numpy(), c.z[f"field_{tag}_{k}"]) < 2e-6, (tag, k)
    # the renderer is network-agnostic (nerf_render.py:149-187): composite of the golden field values
    comp = orc.composite(c.t("dists_fine"), c.t("field_fine_density"), c.t("field_fine_color"), c.rc.max_dist)
    for k in ("color", "depth", "transmittance"):
        assert nerr(comp[k].numpy(), c.z["out_" + k]) < 2e-6, k
