#!/usr/bin/env python
"""Golden fixtures for the NeRF variant's TRAINING backward, from the REAL reference's autograd.

Run in the build container only (needs /root/reference):

    python tests/golden/make_nerf_train_golden.py

The reference's NeRF (neddf/network/nerf.py) inside its NeRFRender, grad mode, recorded uniforms; loss =
color.sum() + 0.1 depth.sum() + 0.05 transmittance.sum() + 0.1 color_coarse.sum() + 0.02 depth_coarse.sum().
Stored in case_nerf_train_<name>.npz: configuration, weights, inputs, the integrated outputs, the fine edge distances,
the upstream gradients that reach the field outputs of each pass (d loss / d density [B,S], d loss / d color [B,S,3],
captured with tensor hooks) and the parameter gradients (every 8th output row of the big matrices).
  relu     default NeRF (ReLU / ReLU, skip 4), one network for both passes, cone sampling, eval low-pass state
  tanhexp  tanhExp hidden, 7 layers, skips [2, 5], separate coarse network, point sampling, low-pass warm-up active
           (set_iter(1500)); ReLU density (see main: the reference's backward raises with negative coarse weights)
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402
import make_nerf_golden as mn  # noqa: E402

import numpy as np  # noqa: E402
import torch  # noqa: E402

RAYS = {"relu": 4, "tanhexp": 6}  # a 64-sample tile costs the host emulation ~0.7 s: small fixtures keep the CPU suite short


def main():
    for name, c in mn.CASES.items():
        # the reference cannot back-propagate through negative coarse weights: sample_pdf zeroes them IN PLACE
        # (base_neural_render.py:52-55) after integrate_volume_render saved them for backward, so autograd raises
        # ("modified by an inplace operation") with the LeakyReLU density of this case; the training fixture uses ReLU
        # (the LeakyReLU slope of the backward kernel is checked against the oracle's autograd instead)
        c = dict(c, net=dict(c["net"], density_activation_type="ReLU"))
        torch.manual_seed(c["seed"] + 100)
        render = mg.build_render(c["net"], c["render"])
        with torch.no_grad():
            for net in {id(n): n for n in (render.network_coarse, render.network_fine)}.values():
                net.outL_density.weight.mul_(8.0)
                net.outL_density.bias.add_(0.5)
        cam = mg.synthetic_camera(c["seed"])
        g = torch.Generator().manual_seed(c["seed"] + 100)
        B = RAYS[name]
        uv = torch.stack([torch.randint(250, 550, (B,), generator=g), torch.randint(250, 550, (B,), generator=g)], 1)
        u_c = torch.rand(B, render.sample_coarse + 1, generator=g)
        u_f = torch.rand(B, render.sample_fine + 1, generator=g)
        render.set_iter(c["iter"])
        fields, ups, hooks, seen = [], [], [], set()

        def fwd_hook(m, i, o):
            fields.append({k: v.detach().clone() for k, v in o.items()})
            slot = {}
            ups.append(slot)
            for k in ("density", "color"):
                o[k].register_hook(lambda gr, k=k, slot=slot: slot.__setitem__(k, gr.detach().clone()))

        for net in (render.network_coarse, render.network_fine):
            if id(net) not in seen:
                seen.add(id(net))
                hooks.append(net.register_forward_hook(fwd_hook))
        pdf_out = []
        orig_pdf = render.sample_pdf
        render.sample_pdf = lambda *a, **k: (pdf_out.append(orig_pdf(*a, **k).detach().clone()) or pdf_out[-1])
        with mg.RandFeeder([u_c, u_f]):
            with torch.set_grad_enabled(True):
                out = render.render_rays(uv, cam)
        render.sample_pdf = orig_pdf
        for h in hooks:
            h.remove()
        loss = (out["color"].sum() + 0.1 * out["depth"].sum() + 0.05 * out["transmittance"].sum()
                + 0.1 * out["color_coarse"].sum() + 0.02 * out["depth_coarse"].sum())
        render.zero_grad()
        loss.backward()
        res = dict(uv=uv.numpy(), u_coarse=u_c.numpy(), u_fine=u_f.numpy(), loss=loss.detach().numpy(), dists_fine=pdf_out[0].numpy(),
                   **mg.cam_arrays(cam))
        for k, v in out.items():
            res["out_" + k] = v.detach().numpy()
        for tag, f, u in zip(("coarse", "fine"), fields, ups):
            for k, v in f.items():
                res[f"field_{tag}_{k}"] = v.numpy()
            for k, v in u.items():
                res[f"up_{tag}_{k}"] = v.numpy()
        for n, p in render.named_parameters():
            gr = p.grad.detach().numpy()
            res["grad_" + n] = gr[::8] if (gr.ndim == 2 and gr.shape[0] > 3) else gr
        res["cfg"] = json.dumps({"net": c["net"], "render": c["render"], "iter": c["iter"], "seed": c["seed"]})
        nets = [("fine", render.network_fine)]
        if render.network_coarse is not render.network_fine:
            nets.append(("coarse", render.network_coarse))
        for tag, net in nets:
            for k, v in net.state_dict().items():
                res[f"w_{tag}.{k}"] = v.detach().numpy()
        np.savez_compressed(os.path.join(HERE, f"case_nerf_train_{name}.npz"), **res)
        print(name, float(loss), {k: getattr(v, "shape", None) for k, v in res.items() if k.startswith(("up_", "grad_"))})


if __name__ == "__main__":
    main()
