"""NeuS field variant: the oracle restatement (oracle.neus_forward, neus.py:101-162) against
goldens recorded from the REAL reference (tests/golden/make_neus_golden.py: the reference's NeuS inside its
NeRFRender in grad mode, recorded uniforms) - per-sample sdf / density / colour and the composited render - and
the forward-mode statement of the same network (the formulation of the CUDA kernel) against the autograd one."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import neddf_oracle as orc
from tests.helpers import GOLDEN, nerr


class NeusCase:
    def __init__(self, name: str):
        z = np.load(os.path.join(GOLDEN, f"case_neus_{name}.npz"), allow_pickle=False)
        self.z = {k: z[k] for k in z.files}
        meta = json.loads(str(self.z["cfg"]))
        self.net_cfg, self.render_cfg = meta["net"], meta["render"]
        self.nc = orc.NeusConfig.from_dict(self.net_cfg)
        self.rc = orc.RenderConfig.from_dict(self.render_cfg)
        cal = [float(v) for v in self.z["cam_calib"]]
        self.cam = orc.CameraPose(torch.from_numpy(self.z["cam_R"]), torch.from_numpy(self.z["cam_T"]), *cal)

    def params(self, tag: str, dtype=torch.float32):
        """[in,out] weights under the reference's state_dict names (torch Linear stores [out,in]) + variance."""
        pre = f"w_{tag}." if f"w_{tag}.layers_sdf.0.weight" in self.z else "w_fine."
        out = {}
        for k, v in self.z.items():
            if k.startswith(pre):
                t = torch.from_numpy(v).to(dtype)
                out[k[len(pre):]] = t.t().contiguous() if k.endswith(".weight") else t
        return out

    def t(self, k):
        return torch.from_numpy(self.z[k])


@pytest.mark.parametrize("name", ["relu", "tanhexp"])
def test_neus_oracle_matches_reference(name):
    c = NeusCase(name)
    shapes = {n: (i, o) for n, i, o in orc.neus_layer_shapes(c.nc)}
    pf = c.params("fine")
    assert {k[:-7] for k in pf if k.endswith(".weight")} == set(shapes)
    assert all(tuple(pf[n + ".weight"].shape) == shapes[n] for n in shapes)
    d, o = orc.make_rays(c.t("uv"), c.cam)
    for tag, dists in (("coarse", orc.coarse_dists(c.rc, c.t("u_coarse"))), ("fine", c.t("dists_fine"))):
        pos, dd, _ = orc.make_samples(c.rc, d, o, dists)
        out = orc.neus_forward(c.params(tag), c.nc, pos, dd)
        fwd = orc.neus_forward_jac(c.params(tag), c.nc, pos, dd)
        for k in ("sdf", "density", "color"):
            assert nerr(out[k].numpy(), c.z[f"field_{tag}_{k}"]) < 2e-6, (tag, k)
            # forward-mode Jacobian rows instead of the reverse-mode gradient: same numbers up to fp32 rounding
            assert nerr(fwd[k].numpy(), c.z[f"field_{tag}_{k}"]) < 2e-5, (tag, k, "forward mode")
        assert nerr(fwd["gradients"].numpy(), out["gradients"].numpy()) < 2e-5
    # the renderer is network-agnostic (nerf_render.py:149-187): composite of the golden field values
    comp = orc.composite(c.t("dists_fine"), c.t("field_fine_density"), c.t("field_fine_color"), c.rc.max_dist)
    for k in ("color", "depth", "transmittance"):
        assert nerr(comp[k].numpy(), c.z["out_" + k]) < 2e-6, k
    assert sorted(k[4:] for k in c.z if k.startswith("out_")) == sorted(
        [a + b for a in ("weight", "depth", "color", "transmittance") for b in ("", "_coarse")])  # no penalty keys


def test_neus_forward_mode_equals_autograd_in_fp64():
    for act in ("ReLU", "tanhExp"):
        cfg = orc.NeusConfig(activation_type=act, sdf_layer_count=5, col_layer_count=3, skips=[1, 3])
        P = {k: v.double() for k, v in orc.neus_init_params(cfg, 3).items()}
        g = torch.Generator().manual_seed(0)
        pos = torch.randn(3, 9, 3, generator=g).double()
        dd = torch.nn.functional.normalize(torch.randn(3, 9, 3, generator=g), dim=-1).double()
        a, b = orc.neus_forward(P, cfg, pos, dd), orc.neus_forward_jac(P, cfg, pos, dd)
        for k in a:
            assert float((a[k] - b[k]).abs().max()) < 1e-12, (act, k)


def test_relu_normal_outliers_sit_on_kinks():
    """Why the GPU test of the normal exempts a few ReLU samples.  The reference restatement itself, in fp32, fed
    positions shifted by ONE ulp: sdf, density and colour move by 1e-6, but the normal of a few samples moves by
    1e-4 .. 1e-3 - exactly the samples where the fp64 run has a hidden pre-activation within rounding distance of
    zero (the ReLU slope, hence the reverse-mode gradient of neus.py:133-142, flips).  Smooth tanhExp: none."""
    for name, expect_outliers in (("relu", True), ("tanhexp", False)):
        c = NeusCase(name)
        d, o = orc.make_rays(c.t("uv"), c.cam)
        pos, dd, _ = orc.make_samples(c.rc, d, o, c.t("dists_fine"))
        P = c.params("fine")
        base = orc.neus_forward(P, c.nc, pos, dd)
        kink = orc.neus_kink_distance(c.params("fine", torch.float64), c.nc, pos.double()).numpy()
        n_out = 0
        for sign in (1.0, -1.0):
            moved = orc.neus_forward(P, c.nc, torch.nextafter(pos, torch.full_like(pos, sign * float("inf"))), dd)
            for k in ("sdf", "density", "color"):
                assert nerr(moved[k].numpy(), base[k].numpy()) < 1e-5, (name, k)
            err = (moved["gradients"] - base["gradients"]).abs().amax(-1).numpy() / float(base["gradients"].abs().max())
            bad = np.argwhere(err >= 1e-4)
            n_out += len(bad)
            assert len(bad) <= max(2, err.size // 500) and float(err.max()) < 5e-2
            assert all(kink[tuple(i)] < 5e-6 for i in bad), [(tuple(i), err[tuple(i)], kink[tuple(i)]) for i in bad]
        assert (n_out > 0) == expect_outliers, (name, n_out)
