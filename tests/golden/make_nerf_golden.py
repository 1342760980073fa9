#!/usr/bin/env python
"""Golden fixtures for the NeRF field variant, from the REAL reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_nerf_golden.py

Drives the reference's own NeRF network (neddf/network/nerf.py) inside its NeRFRender
(neddf/render/nerf_render.py) with recorded uniforms, like make_golden.py does for NeDDF, and stores
weights, inputs and outputs in case_nerf_<name>.npz:
  relu     default construction (ReLU / ReLU, skip 4), one network for both passes, cone sampling, eval state
  tanhexp  tanhExp hidden + LeakyReLU density, 7 layers, skips [2, 5], separate coarse network, point sampling,
           low-pass warm-up active (set_iter(1500))
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
        net={"_target_": "neddf.network.NeRF", "embed_pos_rank": 10, "embed_dir_rank": 4, "layer_count": 8,
             "layer_width": 256, "activation_type": "ReLU", "density_activation_type": "ReLU", "skips": [4],
             "lowpass_alpha_offset": 10.0},
        render={"sample_coarse": 64, "sample_fine": 128, "dist_near": 2.0, "dist_far": 6.0, "max_dist": 6.0,
                "use_coarse_network": False, "sampling_type": "cone"},
        iter=-1, seed=11, rays=48),
    "tanhexp": dict(
        net={"_target_": "neddf.network.NeRF", "embed_pos_rank": 8, "embed_dir_rank": 3, "layer_count": 7,
             "layer_width": 256, "activation_type": "tanhExp", "density_activation_type": "LeakyReLU",
             "skips": [2, 5], "lowpass_alpha_offset": 4.0},
        render={"sample_coarse": 32, "sample_fine": 48, "dist_near": 1.5, "dist_far": 5.0, "max_dist": 5.5,
                "use_coarse_network": True, "sampling_type": "point"},
        iter=1500, seed=12, rays=40),
}


def main():
    for name, c in CASES.items():
        torch.manual_seed(c["seed"])
        render = mg.build_render(c["net"], c["render"])
        # torch's default Linear init is small for a 60-d positional encoding: scale the density head up so that
        # the composited weights are not all ~0 (the fixture carries its weights, any values are legitimate)
        with torch.no_grad():
            for net in {id(n): n for n in (render.network_coarse, render.network_fine)}.values():
                net.outL_density.weight.mul_(8.0)
                net.outL_density.bias.add_(0.5)
        cam = mg.synthetic_camera(c["seed"])
        g = torch.Generator().manual_seed(c["seed"])
        uv = torch.stack([torch.randint(250, 550, (c["rays"],), generator=g), torch.randint(250, 550, (c["rays"],), generator=g)], 1)
        arrays = mg.run_case(render, cam, uv, c["seed"], c["iter"])
        out = dict(arrays)
        out.update(mg.cam_arrays(cam))
        out["uv"] = uv.numpy()
        out["cfg"] = json.dumps({"net": c["net"], "render": c["render"], "iter": c["iter"], "seed": c["seed"]})
        nets = [("fine", render.network_fine)]
        if render.network_coarse is not render.network_fine:
            nets.append(("coarse", render.network_coarse))
        for tag, net in nets:
            for k, v in net.state_dict().items():
                out[f"w_{tag}.{k}"] = v.detach().numpy()
        np.savez_compressed(os.path.join(HERE, f"case_nerf_{name}.npz"), **out)
        print(name, {k: getattr(v, "shape", None) for k, v in out.items() if not k.startswith("w_")})


if __name__ == "__main__":
    main()
