#!/usr/bin/env python
"""Generate the golden parity fixtures by running the REAL reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

It imports ueda0319/neddf from /root/reference (with the hydra/omegaconf stand-ins in
tests/golden/_refstub, neither is installed here), drives the reference's own
NeRFRender / NeDDF / Camera classes on fixed inputs and stores inputs + outputs
as .npz next to this file.  torch.rand inside the reference (nerf_render.py:137,
base_neural_render.py:75) is patched so the uniforms are the recorded ones.

Cases (see CASES below):
  bunny    pretrained bunny_smoke weights + its saved hydra config, real test pose 0
  default  config/network/neddf.yaml + config/render/neddf_render.yaml, seeded weights,
           also a training-state run (set_iter(3000)) with parameter gradients
  point    (historic name) ReLU hidden, tanhExp density, separate coarse network, low-pass
           warm-up active, non-default sample counts / near / far / max_dist
  leaky    LeakyReLU hidden + density, cone sampling
  image    bunny test frame 0 rendered at downsampling=10 through render_image
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path[:0] = [REF, os.path.join(HERE, "_refstub"), REPO]

import cv2  # noqa: E402
import numpy as np  # noqa: E402
import torch  # noqa: E402
import yaml  # noqa: E402

from neddf.camera import Camera, PinholeCalib  # noqa: E402  (reference)
from neddf.render import NeRFRender  # noqa: E402  (reference)
from scipy.spatial.transform import Rotation  # noqa: E402

from oracle import neddf_oracle as orc  # noqa: E402  (only for the seeded weight init)

torch.set_num_threads(8)


class RandFeeder:
    """Replaces torch.rand inside the reference with pre-drawn uniforms, in call order."""

    def __init__(self, tensors):
        self.q = list(tensors)
        self.orig = torch.rand

    def __enter__(self):
        def fake(*shape, **kw):
            t = self.q.pop(0)
            assert tuple(t.shape) == tuple(shape), (t.shape, shape)
            return t.clone()

        torch.rand = fake
        return self

    def __exit__(self, *a):
        torch.rand = self.orig
        assert not self.q


def build_render(net_cfg, render_cfg):
    r = {k: v for k, v in render_cfg.items() if k != "_target_"}
    return NeRFRender(network_config=dict(net_cfg), **r)


def bunny_camera(frame=0, split="test"):
    tf = json.load(open(f"{REF}/data/bunny_smoke/transforms_{split}.json"))
    img = cv2.imread(f"{REF}/data/bunny_smoke/{tf['frames'][frame]['file_path']}.png", cv2.IMREAD_UNCHANGED)
    h, w = img.shape[:2]
    focal = 0.5 * w / np.tan(0.5 * float(tf["camera_angle_x"]))
    m = np.array(tf["frames"][frame]["transform_matrix"])
    p = np.zeros(6, np.float32)
    p[:3] = Rotation.from_matrix(m[:3, :3]).as_rotvec()
    p[3:] = m[:3, 3]
    calib = PinholeCalib(np.array([focal, focal, 0.5 * w, 0.5 * h]))
    cam = Camera(calib, p)
    cam.update_transform()
    rgb = (1.0 / 256) * img[:, :, 3, None].astype(np.float32) * img[:, :, :3].astype(np.float32)
    return cam, w, h, rgb.astype(np.uint8)


def synthetic_camera(seed, w=800, h=800, radius=4.0311):
    """Pose on a sphere looking at the origin; goes through the reference Camera."""
    g = np.random.default_rng(seed)
    v = g.normal(size=3)
    v /= np.linalg.norm(v)
    pos = radius * v
    back = v  # camera looks along -z (RUB), so +z axis points away from the origin
    up = np.array([0.0, 0.0, 1.0])
    right = np.cross(up, back)
    right /= np.linalg.norm(right)
    up2 = np.cross(back, right)
    Rm = np.stack([right, up2, back], 1)
    p = np.zeros(6, np.float32)
    p[:3] = Rotation.from_matrix(Rm).as_rotvec()
    p[3:] = pos
    focal = 0.5 * w / np.tan(0.5 * 0.6911112070083618)
    cam = Camera(PinholeCalib(np.array([focal, focal, 0.5 * w, 0.5 * h])), p)
    cam.update_transform()
    return cam


def cam_arrays(cam):
    prm = cam.camera_calib.params.detach().numpy()
    return dict(cam_R=cam.R.detach().numpy(), cam_T=cam.T.detach().numpy(), cam_calib=prm.astype(np.float32))


def run_case(render, cam, uv, seed, it, with_grad=False):
    """One render_rays call on the reference with recorded uniforms; returns arrays to store."""
    B = uv.shape[0]
    g = torch.Generator().manual_seed(seed)
    u_c = torch.rand(B, render.sample_coarse + 1, generator=g)
    u_f = torch.rand(B, render.sample_fine + 1, generator=g)
    render.set_iter(it)
    # capture per-sample field outputs and the fine distances
    fields = []
    hooks = []
    seen = set()
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
    with RandFeeder([u_c, u_f]):
        with torch.set_grad_enabled(with_grad):
            out = render.render_rays(uv, cam)
    render.sample_pdf = orig_pdf
    for h in hooks:
        h.remove()
    res = dict(uv=uv.numpy(), u_coarse=u_c.numpy(), u_fine=u_f.numpy(), iter=np.int64(it), **cam_arrays(cam))
    for k, v in out.items():
        res["out_" + k] = v.detach().numpy()
    for tag, f in zip(("coarse", "fine"), fields):
        for k, v in f.items():
            res[f"field_{tag}_{k}"] = v.numpy()
    res["dists_fine"] = pdf_out[0].numpy()
    if with_grad:
        loss = (out["color"].sum() + 0.1 * out["depth"].sum() + 0.05 * out["transmittance"].sum()
                + 0.01 * out["fields_penalty"].sum() + 0.1 * out["color_coarse"].sum()
                + 0.001 * out["fields_penalty_coarse"].sum())
        render.zero_grad()
        loss.backward()
        res["loss"] = loss.detach().numpy()
        for n, p in render.named_parameters():
            g = p.grad.detach().numpy()
            # big weight matrices: keep every 8th input row to keep the fixture small
            res["grad_" + n] = g[::8] if (g.ndim == 2 and g.shape[1] > 3) else g
    return res


def pick_uv(w, h, n, seed, centre_frac=0.7):
    g = np.random.default_rng(seed)
    k = int(n * centre_frac)
    cu = g.integers(int(0.3 * w), int(0.7 * w), size=k)
    cv_ = g.integers(int(0.3 * h), int(0.7 * h), size=k)
    eu = g.integers(0, w, size=n - k)
    ev = g.integers(0, h, size=n - k)
    uv = np.stack([np.concatenate([cu, eu]), np.concatenate([cv_, ev])], 1)
    uv[-1] = (0, 0)
    uv[-2] = (w - 1, h - 1)
    return torch.from_numpy(uv.astype(np.int64))


def load_seeded(render, cfg_net, seed, bias_std, separate):
    fc = orc.FieldConfig.from_dict(cfg_net)
    pf = orc.init_params(fc, seed, bias_std)
    sd = {"network_fine." + k: v for k, v in pf.items()}
    pc = orc.init_params(fc, seed + 1, bias_std) if separate else pf
    sd.update({"network_coarse." + k: v for k, v in pc.items()})
    print("   load:", render.load_state_dict(sd))


def main():
    out_dir = HERE
    # ---------------- bunny: pretrained --------------------------------------------------
    cfg = yaml.safe_load(open(f"{REF}/pretrained/bunny_smoke/.hydra/config.yaml"))
    render = build_render(cfg["network"], cfg["render"])
    sd = torch.load(f"{REF}/pretrained/bunny_smoke/models/model_02000.pth", map_location="cpu")
    print("bunny load:", render.load_state_dict(sd))
    # checkpoint weights: coarse and fine alias the same tensors, store the fine copy only
    np.savez(os.path.join(out_dir, "bunny_smoke_weights.npz"),
             **{k[len("network_fine."):]: v.numpy() for k, v in sd.items() if k.startswith("network_fine.")})
    cam, w, h, gt = bunny_camera(0)
    uv = pick_uv(w, h, 24, seed=1)
    res = run_case(render, cam, uv, seed=11, it=-1)
    res["cfg"] = np.array(json.dumps(dict(network=cfg["network"], render=cfg["render"], weights="bunny_smoke_weights.npz")))
    np.savez_compressed(os.path.join(out_dir, "case_bunny.npz"), **res)
    print("bunny done", {k: float(np.abs(v).max()) for k, v in res.items() if k.startswith("out_")})

    # ---------------- image: bunny frame 0 at downsampling 10 ---------------------------
    ds, chunk = 10, 500
    n_pix = (w // ds) * (h // ds)
    g = torch.Generator().manual_seed(123)
    u_c = torch.rand(n_pix, 65, generator=g)
    u_f = torch.rand(n_pix, 129, generator=g)
    feed = []
    for b in range(0, n_pix, chunk):
        feed += [u_c[b:b + chunk], u_f[b:b + chunk]]
    render.set_iter(-1)
    with RandFeeder(feed):
        img = render.render_image(w, h, cam, ["color", "depth", "transmittance"], ds, chunk)
    np.savez_compressed(os.path.join(out_dir, "case_image.npz"), width=w, height=h, downsampling=ds, chunk=chunk,
                        rand_seed=123, color=img["color"].numpy(), depth=img["depth"].numpy(),
                        transmittance=img["transmittance"].numpy(), gt_bgr_u8=gt[::ds, ::ds],
                        cfg=np.array(json.dumps(dict(network=cfg["network"], render=cfg["render"], weights="bunny_smoke_weights.npz"))),
                        **cam_arrays(cam))
    rgb = np.clip(img["color"].numpy() * 255, 0, 255).astype(np.uint8)
    mse = np.mean((rgb.astype(np.float64) - gt[::ds, ::ds].astype(np.float64)) ** 2)
    print("image done, psnr vs gt", 10 * np.log10(255.0 ** 2 / mse))

    # ---------------- default config, seeded weights ------------------------------------
    net = yaml.safe_load(open(f"{REF}/config/network/neddf.yaml"))
    rnd = yaml.safe_load(open(f"{REF}/config/render/neddf_render.yaml"))
    render = build_render(net, rnd)
    load_seeded(render, net, seed=3408, bias_std=0.05, separate=False)
    cam = synthetic_camera(7)
    uv = pick_uv(800, 800, 16, seed=2)
    res = run_case(render, cam, uv, seed=12, it=-1)
    res["cfg"] = np.array(json.dumps(dict(network=net, render=rnd, weights=dict(seed=3408, bias_std=0.05, separate=False))))
    np.savez_compressed(os.path.join(out_dir, "case_default.npz"), **res)
    # training state with gradients (drums config: forward+backward)
    uv = pick_uv(800, 800, 8, seed=3)
    res = run_case(render, cam, uv, seed=13, it=3000, with_grad=True)
    res["cfg"] = np.array(json.dumps(dict(network=net, render=rnd, weights=dict(seed=3408, bias_std=0.05, separate=False))))
    np.savez_compressed(os.path.join(out_dir, "case_train.npz"), **res)
    print("default/train done")

    # ---------------- point sampling / ReLU / separate coarse net / low-pass ------------
    net_p = dict(net, activation_type="ReLU", density_activation_type="tanhExp", lowpass_alpha_offset=4.0)
    # NOTE: the reference itself cannot run NeDDF with sampling_type="point": its
    # get_sampling_points returns expanded (non-contiguous) tensors and NeDDF.forward calls
    # .view(-1, 3) on them (neddf.py:201,210) -> RuntimeError.  Hence cone sampling here.
    rnd_p = dict(rnd, sampling_type="cone", use_coarse_network=True, sample_coarse=32, sample_fine=48,
                 dist_near=1.5, dist_far=5.5, max_dist=7.0)
    render = build_render(net_p, rnd_p)
    load_seeded(render, net_p, seed=77, bias_std=0.05, separate=True)
    cam = synthetic_camera(8)
    uv = pick_uv(800, 800, 16, seed=4)
    res = run_case(render, cam, uv, seed=14, it=2500)
    res["cfg"] = np.array(json.dumps(dict(network=net_p, render=rnd_p, weights=dict(seed=77, bias_std=0.05, separate=True))))
    np.savez_compressed(os.path.join(out_dir, "case_point.npz"), **res)
    print("point done")

    # ---------------- LeakyReLU hidden + density -----------------------------------------
    net_l = dict(net, activation_type="LeakyReLU", density_activation_type="LeakyReLU")
    render = build_render(net_l, rnd)
    load_seeded(render, net_l, seed=99, bias_std=0.05, separate=False)
    cam = synthetic_camera(9)
    uv = pick_uv(800, 800, 8, seed=5)
    res = run_case(render, cam, uv, seed=15, it=-1)
    res["cfg"] = np.array(json.dumps(dict(network=net_l, render=rnd, weights=dict(seed=99, bias_std=0.05, separate=False))))
    np.savez_compressed(os.path.join(out_dir, "case_leaky.npz"), **res)
    print("leaky done")


if __name__ == "__main__":
    main()
