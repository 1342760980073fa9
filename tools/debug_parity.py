"""Scratch: per-key parity report on the GPU (not a test)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import neddf_b200
from oracle import neddf_oracle as orc
from tests.helpers import Case, nerr
import tests.gpu_util as G

eng = sys.argv[1] if len(sys.argv) > 1 else "fp32"
for name in ["bunny", "default", "point", "leaky"]:
    c = Case(name)
    render, cam = G.build_render(c, eng), G.build_camera(c)
    d_ref, o_ref = orc.make_rays(c.t("uv"), c.cam)
    df = c.t("dists_fine")
    pos, dd, var = orc.make_samples(c.rc, d_ref, o_ref, df)
    with torch.no_grad():
        out = render.network_fine(neddf_b200.Sampling(pos.to(G.DEV), dd.contiguous().to(G.DEV), var.to(G.DEV)))
        rd, ro, dfd = d_ref.to(G.DEV).contiguous(), o_ref.contiguous().to(G.DEV), df.to(G.DEV).contiguous()
        out2 = render.network_fine.forward_rays(rd, ro, dfd, c.rc.sampling_type, render._ray_radius, True, True)
        full = render.render_rays(c.t("uv").to(G.DEV), cam, uniforms=(c.t("u_coarse").to(G.DEV), c.t("u_fine").to(G.DEV)))
    torch.cuda.synchronize()
    print(f"== {name} ({eng})")
    for k in ("distance", "density", "color", "fields_penalty", "aux_grad"):
        ref = c.z["field_fine_" + k]
        e1 = np.abs(out[k].cpu().numpy() - ref) / np.abs(ref).max()
        e2 = np.abs(out2[k].cpu().numpy() - ref) / np.abs(ref).max()
        print(f"  field {k:15s} sampling-path max {e1.max():.2e} p99.9 {np.quantile(e1, 0.999):.2e} | rays-path max {e2.max():.2e} p99.9 {np.quantile(e2,0.999):.2e}")
    for k, v in c.outputs().items():
        print(f"  render {k:22s} {nerr(full[k].cpu().numpy(), v):.2e}")
    # sample_pdf detail
    dists = orc.coarse_dists(c.rc, c.t("u_coarse"))
    w = torch.from_numpy(c.z["out_weight_coarse"]).clone()
    u = c.t("u_fine")
    dd_, wd, ud = dists.to(G.DEV), w.to(G.DEV).contiguous(), u.to(G.DEV)
    B, E = dists.shape
    cdf_d = torch.empty(B, E, device=G.DEV)
    from neddf_b200 import _lib as L
    outp = torch.empty(B, E + u.shape[1], device=G.DEV)
    ids = torch.empty(B, u.shape[1], dtype=torch.int64, device=G.DEV)
    L.check(L.lib().neddf_sample_pdf(L.ptr(dd_), L.ptr(wd), L.ptr(ud), B, E, u.shape[1], L.ptr(outp), L.ptr(ids), L.ptr(cdf_d), None, L.stream_ptr(G.DEV)))
    cdf = orc.pdf_cdf(w)
    print("  cdf max abs diff", float((cdf_d.cpu() - cdf).abs().max()), "n differing", int((cdf_d.cpu() != cdf).sum()), "of", cdf.numel())
    ref = orc.sample_pdf(dists, w, u)
    d = (outp.cpu() - ref).abs()
    print("  dists_fine max abs diff", float(d.max()), "ids mismatch", int((ids.cpu() != orc.invert_cdf(dists, cdf, u)[1]).sum()))
