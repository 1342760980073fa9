#!/usr/bin/env python
"""Golden fixtures for the NeuS field variant, from the REAL reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_neus_golden.py

Drives the reference's own NeuS network (neddf/network/neus.py) inside its NeRFRender
(neddf/render/nerf_render.py) with recorded uniforms and autograd ENABLED - the reference takes the SDF normal
with torch.autograd.grad (neus.py:133-142), so its forward only works in grad mode (its render_image, which
disables grad, raises for this network).  Stores weights, inputs, per-sample field outputs (sdf, density, color)
and the composited render in case_neus_<name>.npz:
  relu     config/network/neus.yaml (ReLU, ranks 6 / 4, 8 + 8 layers, skip 4), one network for both passes, cone
           sampling (the SDF trunk ignores the sample variance)
  tanhexp  tanhExp, ranks 5 / 3, 6 SDF layers with skips [1, 3], 3 colour layers, separate coarse network, point
           sampling, other sample counts
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402  (sets sys.path for the reference + stubs)

import numpy as np  # noqa: E402
import torch  # noqa: E402

CASES = {
    "relu": dict(
        net={"_target_": "neddf.network.NeuS", "embed_pos_rank": 6, "embed_dir_rank": 4, "sdf_layer_count": 8,
             "sdf_layer_width": 256, "col_layer_count": 8, "col_layer_width": 256, "init_variance": 0.3,
             "activation_type": "ReLU", "skips": [4]},
        render={"sample_coarse": 64, "sample_fine": 128, "dist_near": 2.0, "dist_far": 6.0, "max_dist": 6.0,
                "use_coarse_network": False, "sampling_type": "cone"},
        seed=21, rays=40),
    "tanhexp": dict(
        net={"_target_": "neddf.network.NeuS", "embed_pos_rank": 5, "embed_dir_rank": 3, "sdf_layer_count": 6,
             "sdf_layer_width": 256, "col_layer_count": 3, "col_layer_width": 256, "init_variance": 0.45,
             "activation_type": "tanhExp", "skips": [1, 3]},
        render={"sample_coarse": 24, "sample_fine": 40, "dist_near": 1.5, "dist_far": 5.0, "max_dist": 5.5,
                "use_coarse_network": True, "sampling_type": "point"},
        seed=22, rays=33),
}


def run_case(render, cam, uv, seed):
    """One render_rays call of the reference in grad mode with recorded uniforms."""
    B = uv.shape[0]
    g = torch.Generator().manual_seed(seed)
    u_c = torch.rand(B, render.sample_coarse + 1, generator=g)
    u_f = torch.rand(B, render.sample_fine + 1, generator=g)
    render.set_iter(-1)
    fields, hooks, seen = [], [], set()
    for net in (render.network_coarse, render.network_fine):
        if id(net) in seen:
            continue
        seen.add(id(net))
        hooks.append(net.register_forward_hook(lambda m, i, o: fields.append({k: v.detach().clone() for k, v in o.items()})))
    pdf_out = []
    orig_pdf = render.sample_pdf

    def pdf_tap(*a, **k):
        r = orig_pdf(*a, **k)
        pdf_out.append(r.detach().clone())
        return r

    render.sample_pdf = pdf_tap
    with mg.RandFeeder([u_c, u_f]):
        with torch.set_grad_enabled(True):
            out = render.render_rays(uv, cam)
    render.sample_pdf = orig_pdf
    for h in hooks:
        h.remove()
    res = dict(uv=uv.numpy(), u_coarse=u_c.numpy(), u_fine=u_f.numpy(), **mg.cam_arrays(cam))
    for k, v in out.items():
        res["out_" + k] = v.detach().numpy()
    for tag, f in zip(("coarse", "fine"), fields):
        for k, v in f.items():
            res[f"field_{tag}_{k}"] = v.numpy()
    res["dists_fine"] = pdf_out[0].numpy()
    return res


def main():
    for name, c in CASES.items():
        torch.manual_seed(c["seed"])
        render = mg.build_render(c["net"], c["render"])
        # torch's default Linear init leaves the SDF channel (channel 0 of the last SDF layer, AFTER the activation)
        # tiny or dead; widen it and lift its bias so that sdf, its gradient and the density vary over the samples
        # (the fixture carries its weights: any values are legitimate)
        with torch.no_grad():
            for net in {id(n): n for n in (render.network_coarse, render.network_fine)}.values():
                last = net.layers_sdf[-1]
                last.weight[0].mul_(6.0)
                last.bias[0].add_(0.35)
                net.layers_col[-1].weight.mul_(3.0)
                net.layers_col[-1].bias.add_(0.3)
        cam = mg.synthetic_camera(c["seed"])
        g = torch.Generator().manual_seed(c["seed"])
        uv = torch.stack([torch.randint(250, 550, (c["rays"],), generator=g), torch.randint(250, 550, (c["rays"],), generator=g)], 1)
        out = run_case(render, cam, uv, c["seed"])
        out["cfg"] = json.dumps({"net": c["net"], "render": c["render"], "seed": c["seed"]})
        nets = [("fine", render.network_fine)]
        if render.network_coarse is not render.network_fine:
            nets.append(("coarse", render.network_coarse))
        for tag, net in nets:
            for k, v in net.state_dict().items():
                out[f"w_{tag}.{k}"] = v.detach().numpy()
        np.savez_compressed(os.path.join(HERE, f"case_neus_{name}.npz"), **out)
        print(name, {k: getattr(v, "shape", None) for k, v in out.items() if not k.startswith("w_")})
        for k in ("field_fine_sdf", "field_fine_density", "field_fine_color", "out_color", "out_transmittance"):
            v = out[k]
            print("   ", k, float(v.min()), float(v.mean()), float(v.max()), "zeros", float((v == 0).mean()))


if __name__ == "__main__":
    main()
