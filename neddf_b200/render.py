"""NeRFRender: the reference's renderer surface over the CUDA hot path.

Reference: neddf/render/nerf_render.py (NeRFRender), neddf/render/base_neural_render.py
(BaseNeuralRender).  Same constructor, parameter names, methods and output dictionary;
ray generation, stratified sampling, the field network, compositing and hierarchical
resampling all run in libneddf_b200.so.

Random numbers: the reference draws torch.rand on the host for the stratified jitter
(nerf_render.py:137) and the inverse-CDF samples (base_neural_render.py:75) - also in eval.
Here the uniforms are kernel *inputs*: by default they are drawn with torch.rand on the
device; pass ``uniforms=(u_coarse[B,S_c+1], u_fine[B,S_f+1])`` to make a call reproducible /
comparable with the reference fed the same numbers.
"""
import importlib
import math
import warnings
from typing import Any, Dict, Iterable, List, Optional, Tuple

import numpy as np
import torch
from torch import Tensor, nn

from . import _lib as L
from .nerf import NeRF
from .network import BaseNeuralField, EngineRangeError, NeDDF
from .neus import NeuS
from .ray import CONE_RAY_RADIUS, Ray

_TARGET_ALIASES = {
    "neddf.network.NeDDF": NeDDF,
    "neddf.network.neddf.NeDDF": NeDDF,
    "neddf_b200.NeDDF": NeDDF,
    "neddf_b200.network.NeDDF": NeDDF,
    # the NeRF field variant: forward / image rendering on the CUDA kernel of nerf.py
    "neddf.network.NeRF": NeRF,
    "neddf.network.nerf.NeRF": NeRF,
    "neddf_b200.NeRF": NeRF,
    "neddf_b200.nerf.NeRF": NeRF,
    # the NeuS field variant (same item): the normal is carried forward inside the kernel, so no-grad renders work
    "neddf.network.NeuS": NeuS,
    "neddf.network.neus.NeuS": NeuS,
    "neddf_b200.NeuS": NeuS,
    "neddf_b200.neus.NeuS": NeuS,
}


def _instantiate(network_config) -> BaseNeuralField:
    """hydra.utils.instantiate(network_config) (nerf_render.py:67-73) without requiring hydra:
    ``_target_`` strings that name the reference's NeDDF resolve to the CUDA-backed class."""
    cfg = {k: network_config[k] for k in network_config.keys()}
    target = cfg.pop("_target_", "neddf.network.NeDDF")
    cls = _TARGET_ALIASES.get(target)
    if cls is None:
        raise NotImplementedError(
            f"neddf_b200.NeRFRender runs the NeDDF field and the NeRF / NeuS variants; network _target_={target!r} "
            "is not one of them - use the reference renderer for it")
    for k in ("skips",):
        if cfg.get(k) is not None:
            cfg[k] = [int(s) for s in cfg[k]]
    if cfg.get("penalty_weight") is not None:
        cfg["penalty_weight"] = {k: float(v) for k, v in dict(cfg["penalty_weight"]).items()}
    return cls(**cfg)


def _camera_host(camera) -> Tuple[Any, Any, Any]:
    """R[9], T[3], calib[4] as ctypes float arrays.  The three tensors are read back ONCE per camera state:
    the result is cached on the camera object and keyed by the tensors' identity and version counters
    (Camera.update_transform rebinds R / T; in-place edits bump _version), so a training loop that renders
    from the same pose pays no device synchronisation per step."""
    R_t, T_t, C_t = camera.R, camera.T, camera.camera_calib.params
    key = tuple((id(x), x.data_ptr(), x._version) for x in (R_t, T_t, C_t))
    cached = getattr(camera, "_neddf_b200_host", None)
    if cached is not None and cached[0] == key:
        return cached[1]
    flat = torch.cat([R_t.detach().reshape(-1).to(torch.float32), T_t.detach().reshape(-1).to(torch.float32),
                      C_t.detach().reshape(-1).to(torch.float32)]).to("cpu").tolist()  # one D2H copy
    nR, nT = R_t.numel(), T_t.numel()
    if nR != 9 or nT != 3 or len(flat) - 12 < 4:
        raise ValueError("camera must expose R[3,3], T[3] and camera_calib.params=[fx,fy,cx,cy]")
    out = (L.fbuf(flat[:9]), L.fbuf(flat[9:12]), L.fbuf(flat[12:16]))
    try:
        camera._neddf_b200_host = (key, out)
    except Exception:  # objects that refuse new attributes just pay the copy every time
        pass
    return out


class _CompositeFn(torch.autograd.Function):
    """Differentiable compositing: forward neddf_composite, backward neddf_composite_backward
    (gradients w.r.t. densities, colours and penalties; edge distances carry none, like the
    reference where they come from torch.rand / a no_grad resampling)."""

    @staticmethod
    def forward(ctx, dists, densities, colors, penalties, max_dist, status):
        B, E = dists.shape
        device = dists.device
        weight = torch.empty(B, E - 1, device=device, dtype=torch.float32)
        depth = torch.empty(B, device=device, dtype=torch.float32)
        color = torch.empty(B, 3, device=device, dtype=torch.float32)
        trans = torch.empty(B, device=device, dtype=torch.float32)
        pen = torch.empty(B, device=device, dtype=torch.float32) if penalties is not None else None
        with torch.cuda.device(device):
            L.check(L.lib().neddf_composite(L.ptr(dists), L.ptr(densities), L.ptr(colors), L.ptr(penalties), B, E,
                                            float(max_dist), L.ptr(weight), L.ptr(depth), L.ptr(color), L.ptr(trans),
                                            L.ptr(pen), L.ptr(status), L.stream_ptr(device)), "composite")
        ctx.save_for_backward(dists, densities, colors)
        ctx.max_dist = float(max_dist)
        ctx.has_pen = penalties is not None
        if pen is None:
            pen = torch.zeros(0, device=device)
        return weight, depth, color, trans, pen

    @staticmethod
    def backward(ctx, g_w, g_d, g_c, g_t, g_p):
        dists, densities, colors = ctx.saved_tensors
        B, E = dists.shape
        device = dists.device

        def prep(g):
            return None if g is None else g.contiguous().to(torch.float32)

        g_w, g_d, g_c, g_t = prep(g_w), prep(g_d), prep(g_c), prep(g_t)
        g_p = prep(g_p) if ctx.has_pen else None
        d_dens = torch.empty_like(densities)
        d_col = torch.empty_like(colors)
        d_pen = torch.empty_like(densities) if ctx.has_pen else None
        with torch.cuda.device(device):
            L.check(L.lib().neddf_composite_backward(L.ptr(dists), L.ptr(densities), L.ptr(colors), B, E, ctx.max_dist,
                                                     L.ptr(g_w), L.ptr(g_d), L.ptr(g_c), L.ptr(g_t), L.ptr(g_p),
                                                     L.ptr(d_dens), L.ptr(d_col), L.ptr(d_pen), L.stream_ptr(device)),
                    "composite_backward")
        return None, d_dens, d_col, d_pen, None, None


class BaseNeuralRender(nn.Module):
    """neddf/render/base_neural_render.py:11-194 (CUDA-backed sample_pdf / integrate_volume_render)."""

    def __init__(self) -> None:
        super().__init__()
        self.iteration: int = -1

    def set_iter(self, iter: int) -> None:
        self.iteration = iter

    def next_iter(self) -> None:
        self.set_iter(self.iteration + 1)

    # ---- a17 ---------------------------------------------------------------------------
    def sample_pdf(self, dists: Tensor, weights: Tensor, samples_fine: int, cat_coarse: bool = True,
                   uniform_rands: Optional[Tensor] = None, return_ids: bool = False):
        """Hierarchical resampling (base_neural_render.py:27-115), both modes.  ``weights`` is sanitised in
        place like the reference's argument."""
        dists = L.require_cuda_f32(dists, "dists")
        if weights.dtype != torch.float32 or not weights.is_contiguous() or not weights.is_cuda:
            raise RuntimeError("neddf_b200: `weights` must be a contiguous fp32 CUDA tensor (it is updated in place)")
        B, E = dists.shape
        if weights.shape != (B, E - 1):
            raise ValueError(f"weights must be [batch, {E - 1}]")
        device = dists.device
        if uniform_rands is None:
            uniform_rands = torch.rand(B, samples_fine, device=device)
        u = L.require_cuda_f32(uniform_rands, "uniform_rands")
        if u.shape != (B, samples_fine):
            raise ValueError("uniform_rands must be [batch, samples_fine]")
        out = torch.empty(B, (E if cat_coarse else 0) + samples_fine, device=device, dtype=torch.float32)
        ids = torch.empty(B, samples_fine, device=device, dtype=torch.int64) if return_ids else None
        status = self._status(device)
        with torch.cuda.device(device):
            L.check(L.lib().neddf_sample_pdf(L.ptr(dists), L.ptr(weights), L.ptr(u), B, E, samples_fine, 1 if cat_coarse else 0, L.ptr(out),
                                             L.ptr(ids), None, L.ptr(status), L.stream_ptr(device)), "sample_pdf")
        return (out, ids) if return_ids else out

    # ---- a15 ---------------------------------------------------------------------------
    def integrate_volume_render(self, dists: Tensor, densities: Tensor, colors: Tensor,
                                penalties: Optional[Tensor] = None) -> Dict[str, Tensor]:
        """Alpha compositing (base_neural_render.py:117-172); with ``penalties`` also the
        per-ray penalty integral of render_rays (nerf_render.py:153-159)."""
        dists = L.require_cuda_f32(dists, "dists")
        densities = L.require_cuda_f32(densities, "densities")
        colors = L.require_cuda_f32(colors, "colors")
        B, E = dists.shape
        device = dists.device
        if torch.is_grad_enabled() and (densities.requires_grad or colors.requires_grad or
                                        (penalties is not None and penalties.requires_grad)):
            if penalties is not None:
                penalties = L.require_cuda_f32(penalties, "penalties")
            w, d, c, t, p_ = _CompositeFn.apply(dists.detach(), densities, colors, penalties, self.max_dist,
                                                self._status(device))
            res = {"weight": w, "depth": d, "color": c, "transmittance": t}
            if penalties is not None:
                res["fields_penalty"] = p_
            return res
        res = {
            "weight": torch.empty(B, E - 1, device=device, dtype=torch.float32),
            "depth": torch.empty(B, device=device, dtype=torch.float32),
            "color": torch.empty(B, 3, device=device, dtype=torch.float32),
            "transmittance": torch.empty(B, device=device, dtype=torch.float32),
        }
        pen_out = None
        if penalties is not None:
            penalties = L.require_cuda_f32(penalties, "penalties")
            pen_out = torch.empty(B, device=device, dtype=torch.float32)
        status = self._status(device)
        with torch.cuda.device(device):
            L.check(L.lib().neddf_composite(L.ptr(dists), L.ptr(densities), L.ptr(colors), L.ptr(penalties), B, E,
                                            float(self.max_dist), L.ptr(res["weight"]), L.ptr(res["depth"]),
                                            L.ptr(res["color"]), L.ptr(res["transmittance"]), L.ptr(pen_out),
                                            L.ptr(status), L.stream_ptr(device)), "composite")
        if pen_out is not None:
            res["fields_penalty"] = pen_out
        return res

    def _status(self, device) -> Tensor:
        st = getattr(self, "_status_buf", None)
        if st is None or st.device != device:
            # [0] persistent flags read by check_status, [1] per-launch scratch of neddf_sample_pdf
            st = torch.zeros(2, device=device, dtype=torch.int32)
            self._status_buf = st
        return st

    def check_status(self) -> None:
        """The reference asserts `not any(isnan(w))` inside integrate_volume_render
        (base_neural_render.py:155) - a host sync per call.  The kernels record the condition in
        a device flag instead; this reads it (one sync) and raises like the reference."""
        range_error = None
        for net in {id(n): n for n in (getattr(self, "network_coarse", None), getattr(self, "network_fine", None))
                    if n is not None}.values():
            if hasattr(net, "check_engine_status"):
                try:
                    net.check_engine_status()
                except EngineRangeError as e:  # keep going: the other network's flag must be read (and cleared) too
                    range_error = e
        st = getattr(self, "_status_buf", None)
        if range_error is not None:
            if st is not None:
                st.zero_()  # flags of the invalid run
            raise range_error
        if st is None:
            return
        v = int(st[0].item())
        st.zero_()
        if v & 1:
            raise AssertionError("NaN in volume-rendering weights (base_neural_render.py:155)")
        if v & 2:
            print("pdf sampling failed")  # base_neural_render.py:106


class NeRFRender(BaseNeuralRender):
    """Drop-in for neddf.render.NeRFRender (neddf/render/nerf_render.py:20-336)."""

    def __init__(
        self,
        network_config,
        sample_coarse: int = 128,
        sample_fine: int = 128,
        dist_near: float = 2.0,
        dist_far: float = 6.0,
        max_dist: float = 6.0,
        use_coarse_network: bool = True,
        sampling_type: str = "point",
    ) -> None:
        super().__init__()
        self.use_coarse_network = bool(use_coarse_network)
        self.network_fine: BaseNeuralField = _instantiate(network_config)
        if use_coarse_network:
            self.network_coarse: BaseNeuralField = _instantiate(network_config)
        else:
            self.network_coarse = self.network_fine
        self.sample_coarse = int(sample_coarse)
        self.sample_fine = int(sample_fine)
        self.dist_near = float(dist_near)
        self.dist_far = float(dist_far)
        self.max_dist = float(max_dist)
        if sampling_type not in L.SAMPLING_IDS:
            raise ValueError(f"unknown sampling_type {sampling_type!r}")
        self.sampling_type = sampling_type
        # nan asserts cost a device sync; the reference pays four per render_rays call
        self.check_nan = True
        # rays per internal launch in render_image (bounds the [rays, samples] work buffers)
        self.image_chunk = 163840  # 0.76 GB of work buffers; a rank's 80,000- or 160,000-ray shard is one launch sequence
        # Early ray termination (BASELINE.json configs[4]; NOT in the reference, so opt-in): in no-grad image
        # renders the fine pass runs in `termination_segments` depth segments and a ray stops being evaluated
        # once its transmittance falls to `transmittance_eps`.  0.0 = off = the reference's behaviour, bit for
        # bit.  Error bound: |d color| <= eps * max|c|, |d depth| <= eps * max_dist, |d transmittance| <= eps.
        self.transmittance_eps = 0.0
        self.termination_segments = 4
        self._term_counters = None

    # ------------------------------------------------------------------ module surface --
    def get_network(self) -> BaseNeuralField:
        return self.network_fine

    def get_parameters_list(self) -> List[Any]:
        if self.use_coarse_network:
            return list(self.network_coarse.parameters()) + list(self.network_fine.parameters())
        return list(self.network_coarse.parameters())

    def set_iter(self, iter: int) -> None:
        super().set_iter(iter)
        self.network_coarse.set_iter(iter)
        self.network_fine.set_iter(iter)

    def set_engine(self, engine: str) -> None:
        """"auto" | "fp32" (CUDA-core fp32 FMA) | "tc" (tcgen05, split-fp16 operands)."""
        if engine not in L.ENGINE_IDS:
            raise ValueError(engine)
        for net in (self.network_fine, self.network_coarse):
            net.engine = engine
            if hasattr(net, "_range_fallback"):
                net._range_fallback = False  # an explicit choice starts afresh

    @property
    def _ray_radius(self) -> float:
        return CONE_RAY_RADIUS if self.sampling_type == "cone" else 0.0

    # ------------------------------------------------------------------- the hot function --
    def _render_core(self, ray_dir: Tensor, ray_orig: Tensor, u_coarse: Tensor, u_fine: Tensor,
                     full: bool) -> Dict[str, Tensor]:
        """render_rays after ray generation (nerf_render.py:130-188).  ``full`` = produce every
        key of the reference dictionary (penalties included); otherwise only what
        color/depth/transmittance images need."""
        lib = L.lib()
        B = ray_dir.shape[0]
        device = ray_dir.device
        Ec, Ef_new = self.sample_coarse + 1, self.sample_fine + 1
        stream = L.stream_ptr(device)
        dists_c = torch.empty(B, Ec, device=device, dtype=torch.float32)
        L.check(lib.neddf_coarse_dists(L.ptr(u_coarse), B, Ec, self.dist_near, self.dist_far, L.ptr(dists_c), stream),
                "coarse_dists")
        vc = self.network_coarse.forward_rays(ray_dir, ray_orig, dists_c, self.sampling_type, self._ray_radius,
                                              need_penalty=full, need_aux=False)
        ic = self.integrate_volume_render(dists_c, vc["density"], vc["color"], vc.get("fields_penalty"))
        dists_f = self.sample_pdf(dists_c, ic["weight"], Ef_new, uniform_rands=u_fine)
        if (not full and self.transmittance_eps > 0.0 and not torch.is_grad_enabled()
                and hasattr(self.network_fine, "forward_rays_segment")):
            vf = self._fine_pass_terminated(ray_dir, ray_orig, dists_f)
        else:
            vf = self.network_fine.forward_rays(ray_dir, ray_orig, dists_f, self.sampling_type, self._ray_radius,
                                                need_penalty=full, need_aux=False)
        out = self.integrate_volume_render(dists_f, vf["density"], vf["color"], vf.get("fields_penalty"))
        for k in list(ic.keys()):
            out[k + "_coarse"] = ic[k]
        return out

    def _fine_pass_terminated(self, ray_dir: Tensor, ray_orig: Tensor, dists_f: Tensor) -> Dict[str, Tensor]:
        """Fine pass with early ray termination: evaluate a depth segment, update every live ray's
        transmittance, keep the rays with T > eps, continue.  Everything stays on the device (the live-ray
        count is read by the next kernel, not by the host).  Samples that are never evaluated keep density 0
        and so drop out of the compositing sum (base_neural_render.py:148-172 with o_j = 0)."""
        lib = L.lib()
        B, E = dists_f.shape
        device = dists_f.device
        density = torch.zeros(B, E, device=device, dtype=torch.float32)
        color = torch.zeros(B, E, 3, device=device, dtype=torch.float32)
        trans = torch.ones(B, device=device, dtype=torch.float32)
        idx = [torch.empty(B, device=device, dtype=torch.int32) for _ in range(2)]
        cnt = [torch.zeros(1, device=device, dtype=torch.int32) for _ in range(2)]
        if self._term_counters is None or self._term_counters.device != device:
            self._term_counters = torch.zeros(2, device=device, dtype=torch.int64)  # executed, nominal
        executed = self._term_counters[0:1]
        self._term_counters[1] += B * E
        K = max(1, min(int(self.termination_segments), E))
        bounds = [round(k * E / K) for k in range(K + 1)]
        cur_idx, cur_n = None, None
        stream = L.stream_ptr(device)
        for k in range(K):
            e0, seg = bounds[k], bounds[k + 1] - bounds[k]
            self.network_fine.forward_rays_segment(ray_dir, ray_orig, dists_f, self.sampling_type, self._ray_radius, e0, seg,
                                                   cur_idx, cur_n, density, color)
            L.check(lib.neddf_terminate_rays(L.ptr(dists_f), L.ptr(density), B, E, e0, seg, L.ptr(cur_idx), L.ptr(cur_n),
                                             L.ptr(trans), float(self.transmittance_eps), L.ptr(idx[k % 2]),
                                             L.ptr(cnt[k % 2]), L.ptr(executed), stream), "terminate_rays")
            cur_idx, cur_n = idx[k % 2], cnt[k % 2]
        return {"density": density, "color": color}

    def termination_stats(self, reset: bool = True) -> Dict[str, int]:
        """MLP evaluations of the fine passes since the last call: executed (after early termination) and
        nominal (what the reference would have run).  One device synchronisation."""
        if self._term_counters is None:
            return {"executed": 0, "nominal": 0}
        e, n = (int(v) for v in self._term_counters.tolist())
        if reset:
            self._term_counters.zero_()
        return {"executed": e, "nominal": n}

    def _uniforms(self, B: int, device, uniforms) -> Tuple[Tensor, Tensor]:
        Ec, Ef = self.sample_coarse + 1, self.sample_fine + 1
        if uniforms is None:
            return (torch.rand(B, Ec, device=device), torch.rand(B, Ef, device=device))
        u_c, u_f = uniforms
        u_c = L.require_cuda_f32(u_c.to(device, non_blocking=True), "u_coarse")
        u_f = L.require_cuda_f32(u_f.to(device, non_blocking=True), "u_fine")
        if u_c.shape != (B, Ec) or u_f.shape != (B, Ef):
            raise ValueError(f"uniforms must be ([{B},{Ec}], [{B},{Ef}])")
        return u_c, u_f

    def render_rays(self, uv: Tensor, camera, uniforms: Optional[Tuple[Tensor, Tensor]] = None) -> Dict[str, Tensor]:
        """uv[B,2] pixel ids -> weight, depth, color, transmittance, fields_penalty (+ *_coarse)
        (nerf_render.py:109-188)."""
        if not uv.is_cuda:
            raise RuntimeError("neddf_b200: `uv` must live on the CUDA device of the camera/renderer")
        if uv.dtype not in L.UV_DTYPES:
            uv = uv.to(torch.int64)
        uv = uv.contiguous()
        B = uv.shape[0]
        device = uv.device
        hR, hT, hC = _camera_host(camera)
        with torch.cuda.device(device):
            ray_dir = torch.empty(B, 3, device=device, dtype=torch.float32)
            ray_orig = torch.empty(B, 3, device=device, dtype=torch.float32)
            L.check(L.lib().neddf_make_rays(L.ptr(uv), L.UV_DTYPES[uv.dtype], B, hR, hT, hC, L.ptr(ray_dir),
                                            L.ptr(ray_orig), L.stream_ptr(device)), "make_rays")
            u_c, u_f = self._uniforms(B, device, uniforms)
            out = self._render_core(ray_dir, ray_orig, u_c, u_f, full=True)
            if self.check_nan:
                try:
                    self.check_status()
                except EngineRangeError as e:  # engine "auto" left fp16 range: same rays, same uniforms, fp32 engine
                    warnings.warn(str(e), RuntimeWarning)
                    out = None  # release the invalid run (under autograd: its saved activations) before the second one
                    out = self._render_core(ray_dir, ray_orig, u_c, u_f, full=True)
                    self.check_status()
        return out

    def create_rays(self, uv: Tensor, camera) -> Ray:
        """Camera.create_rays (camera.py:155-171) on device."""
        uv = uv.contiguous()
        if uv.dtype not in L.UV_DTYPES:
            uv = uv.to(torch.int64)
        B, device = uv.shape[0], uv.device
        hR, hT, hC = _camera_host(camera)
        ray_dir = torch.empty(B, 3, device=device, dtype=torch.float32)
        ray_orig = torch.empty(B, 3, device=device, dtype=torch.float32)
        with torch.cuda.device(device):
            L.check(L.lib().neddf_make_rays(L.ptr(uv), L.UV_DTYPES[uv.dtype], B, hR, hT, hC, L.ptr(ray_dir),
                                            L.ptr(ray_orig), L.stream_ptr(device)), "make_rays")
        return Ray(ray_dir, ray_orig, uv)

    # ------------------------------------------------------------------------ images ------
    def render_pixels(self, width: int, height: int, camera, target_types: Iterable[str], downsampling: int,
                      first: int, count: int, uniforms=None, device=None) -> Dict[str, Tensor]:
        """Rows [first, first+count) of the row-major pixel list of render_image, as flat
        [count, C] tensors.  This is the unit of work that is sharded across GPUs."""
        target_types = list(target_types)
        lib = L.lib()
        device = torch.device(device) if device is not None else self.network_fine.device
        if device.type != "cuda":
            raise RuntimeError(f"neddf_b200.NeRFRender renders on CUDA devices only (the module is on {device}): move it "
                               "with .to('cuda') - the hot path has no CPU implementation")
        hR, hT, hC = _camera_host(camera)
        outs: Dict[str, List[Tensor]] = {k: [] for k in target_types}
        with torch.no_grad(), torch.cuda.device(device):
            for b0 in range(first, first + count, self.image_chunk):
                n = min(self.image_chunk, first + count - b0)
                ray_dir = torch.empty(n, 3, device=device, dtype=torch.float32)
                ray_orig = torch.empty(n, 3, device=device, dtype=torch.float32)
                L.check(lib.neddf_make_image_rays(int(width), int(height), int(downsampling), b0, n, hR, hT, hC,
                                                  L.ptr(ray_dir), L.ptr(ray_orig), L.stream_ptr(device)),
                        "make_image_rays")
                if uniforms is None:
                    u = None
                else:
                    u = (uniforms[0][b0 - first:b0 - first + n], uniforms[1][b0 - first:b0 - first + n])
                u_c, u_f = self._uniforms(n, device, u)
                res = self._render_core(ray_dir, ray_orig, u_c, u_f, full=False)
                for k in target_types:
                    outs[k].append(res[k].reshape(n, -1))
        return {k: (torch.cat(v, 0) if len(v) != 1 else v[0]) for k, v in outs.items()}

    def render_image(self, width: int, height: int, camera, target_types: Iterable[str], downsampling: int = 1,
                     chunk: int = 512, uniforms=None) -> Dict[str, Tensor]:
        """Whole-image render (nerf_render.py:190-249).  ``chunk`` is accepted for signature
        compatibility; the device path sizes its own launches (``self.image_chunk``) since the
        result does not depend on the chunking."""
        target_types = list(target_types)
        w, h = width // downsampling, height // downsampling
        was_training = (self.network_coarse.training, self.network_fine.training)
        self.network_coarse.eval()
        self.network_fine.eval()
        try:
            flat = self.render_pixels(width, height, camera, target_types, downsampling, 0, w * h, uniforms)
        finally:
            # the reference leaves both networks in train mode afterwards (nerf_render.py:247-248)
            self.network_coarse.train(True)
            self.network_fine.train(True)
        del was_training
        if self.check_nan:
            try:
                self.check_status()
            except EngineRangeError as e:  # engine "auto" left fp16 range: render the frame again on the fp32 engine
                warnings.warn(str(e), RuntimeWarning)
                with torch.no_grad():
                    flat = self.render_pixels(width, height, camera, target_types, downsampling, 0, w * h, uniforms)
                self.check_status()
        return {k: v.reshape(h, w, -1) for k, v in flat.items()}

    def render_field_slice(self, slice_t: float = 0.0, render_size: float = 1.1,
                           render_resolution: int = 128) -> Dict[str, np.ndarray]:
        """Field slice visualisation (nerf_render.py:263-336): uint8 BGR images."""
        import cv2

        from .ray import Sampling

        with torch.no_grad():
            device = self.network_fine.device
            lin = torch.linspace(-render_size, render_size, render_resolution, device=device)
            xs = lin.reshape(1, -1).expand(render_resolution, render_resolution)
            ys = -lin.reshape(-1, 1).expand(render_resolution, render_resolution)
            zs = torch.zeros(render_resolution, render_resolution, device=device) + slice_t
            pos = torch.stack([xs, ys, zs], 2).contiguous()
            sdir = torch.zeros_like(pos)
            sdir[:, :, 2] = 1.0
            self.network_fine.train(False)  # nerf_render.py:309-311 toggles the mode around the query
            values = self.network_fine(Sampling(pos, sdir, torch.zeros_like(pos)))
            self.network_fine.train(True)
            scales = {"distance": 256.0, "density": 12.8, "color": 256.0, "aux_grad": 256.0}
            fields: Dict[str, np.ndarray] = {}
            for key, scale in scales.items():
                f = (scale * values[key].reshape(render_resolution, render_resolution, -1)).cpu().numpy()
                if f.shape[2] == 1:
                    fields[key] = cv2.applyColorMap(f.clip(0, 255).astype(np.uint8), cv2.COLORMAP_JET)
                else:
                    fields[key] = f.clip(0, 255).astype(np.uint8)
            return fields
