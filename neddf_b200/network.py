"""NeDDF field network: the reference's module surface over the CUDA megakernel.

Reference: neddf/network/neddf.py (NeDDF), neddf/network/base_neuralfield.py
(BaseNeuralField), neddf/nn_module/with_grad/linear.py:87-133 (LinearGradLayer parameters).

The module owns the same parameters under the same names/shapes as the reference
(``layers_ddf.{i}.weight`` [in,out] ...), so ``state_dict`` files are interchangeable.  All
arithmetic of ``forward`` happens in libneddf_b200.so (neddf_field_forward*); the parameters
are re-packed into kernel layout whenever their version counters change (optimiser steps
update them in place).
"""
import ctypes as C
import math
from typing import Dict, List, Optional

import numpy as np
import torch
from torch import Tensor, nn

from . import _lib as L
from .ray import Sampling


class LinearGradLayer(nn.Module):
    """Parameter holder with the reference's layout: weight [in,out] xavier-normal, zero bias
    (nn_module/with_grad/linear.py:111-116).  The product computes y = xW + b, G = JW inside
    the fused field kernel, never layer by layer."""

    def __init__(self, input_ch: int = 128, output_ch: int = 128) -> None:
        super().__init__()
        self.input_ch, self.output_ch = input_ch, output_ch
        self.weight = nn.Parameter(torch.randn(input_ch, output_ch))
        self.bias = nn.Parameter(torch.randn(output_ch))
        nn.init.xavier_normal_(self.weight)
        nn.init.constant_(self.bias, 0.0)

    def forward(self, *a, **k):  # pragma: no cover
        raise NotImplementedError("neddf_b200 fuses all linear layers into the field megakernel; "
                                  "call NeDDF.forward instead")


class _FieldTrainFn(torch.autograd.Function):
    """Differentiable NeDDF.forward on rays + edge distances (training path).

    forward : neddf_field_forward_train - the fused megakernel of the module's engine (``net.engine``:
              tensor-core by default, "fp32" on request), keeping every layer's pre-activations.
    backward: neddf_field_backward does all sample-local work (activation second derivatives, heads,
              density, penalties, data-gradient GEMMs) and writes per-layer inputs X_l and
              pre-activation gradients G_l; the weight gradients gW_l = X_l^T G_l are tensor-core split-K
              GEMMs of this library (neddf_wgrad, csrc/wgrad.cu), bias gradients column sums
              (neddf_colsum_value_rows).
    Gradients flow to the module's parameters only (sample positions come from torch.rand / a
    no_grad resampling in the reference, nerf_render.py:131-166).
    """

    @staticmethod
    def forward(ctx, net, a, b, c, sampling_type, ray_radius, *params):
        """(a, b, c) = (ray_dir[B,3], ray_orig[B,3], dists[B,S]) with a sampling type, or the Sampling
        tensors (pos, dir, var)[B,S,3] when sampling_type is None."""
        from_rays = sampling_type is not None
        B, S = (c.shape if from_rays else a.shape[:2])
        n = B * S
        device = a.device
        n_hidden = (net.ddf_layer_count - 1) + (net.col_layer_count - 1)
        h = net._field(device)
        st = net._state_struct()
        save = torch.empty(n_hidden, n, 4, 256, device=device, dtype=torch.float32)
        density = torch.empty(B, S, device=device, dtype=torch.float32)
        color = torch.empty(B, S, 3, device=device, dtype=torch.float32)
        penalty = torch.empty(B, S, device=device, dtype=torch.float32)
        distance = torch.empty(B, S, device=device, dtype=torch.float32) if not from_rays else None
        aux = torch.empty(B, S, device=device, dtype=torch.float32) if not from_rays else None
        with torch.cuda.device(device):
            if from_rays:
                L.check(L.lib().neddf_field_forward_train(
                    h, C.byref(st), L.ptr(a), L.ptr(b), L.ptr(c), B, S, L.SAMPLING_IDS[sampling_type],
                    float(ray_radius), L.ptr(density), L.ptr(color), L.ptr(penalty), L.ptr(save),
                    net._engine_id(), L.stream_ptr(device)), "field_forward_train")
            else:
                L.check(L.lib().neddf_field_forward_train_samples(
                    h, C.byref(st), L.ptr(a), L.ptr(b), L.ptr(c), n, L.ptr(distance), L.ptr(density), L.ptr(color),
                    L.ptr(penalty), L.ptr(aux), L.ptr(save), net._engine_id(), L.stream_ptr(device)),
                    "field_forward_train_samples")
        ctx.net = net
        ctx.meta = (sampling_type, float(ray_radius),
                    (st.aux_grad_scale, st.distance_range_max, st.lowpass_alpha, tuple(st.penalty_weight)), (B, S))
        ctx.save_for_backward(a, b, c, save)
        if from_rays:
            return density, color, penalty
        ctx.mark_non_differentiable(distance, aux)
        return density, color, penalty, distance, aux

    @staticmethod
    def backward(ctx, g_density, g_color, g_penalty, *unused):
        net = ctx.net
        ga, gb_, gc_, save = ctx.saved_tensors
        sampling_type, ray_radius, stv, (B, S) = ctx.meta
        from_rays = sampling_type is not None
        n = B * S
        device = ga.device
        n_ddf, n_col = net.ddf_layer_count - 1, net.col_layer_count - 1
        n_hidden = n_ddf + n_col
        n_e0 = 6 * net.embed_pos_rank
        off_h = 6 * (net.embed_pos_rank + net.embed_dir_rank) + 3

        def prep(g, shape):
            if g is None:
                return torch.zeros(shape, device=device, dtype=torch.float32)
            return g.contiguous().to(torch.float32)

        g_density = prep(g_density, (B, S))
        g_color = prep(g_color, (B, S, 3))
        g_penalty = prep(g_penalty, (B, S))
        post = torch.empty(n_hidden, n, 4, 256, device=device, dtype=torch.float32)
        gpre = torch.empty(n_hidden, n, 4, 256, device=device, dtype=torch.float32)
        ghead_da = torch.empty(n, 4, 2, device=device, dtype=torch.float32)
        ghead_col = torch.empty(n, 4, 4, device=device, dtype=torch.float32)
        xes = torch.empty(n, 4, n_e0, device=device, dtype=torch.float32)
        xcol = torch.empty(n, 4, off_h, device=device, dtype=torch.float32)
        h = net._field(device)
        st = L.FieldState(stv[0], stv[1], stv[2], (C.c_float * L.N_PENALTY)(*stv[3]))
        with torch.cuda.device(device):
            if from_rays:
                L.check(L.lib().neddf_field_backward(
                    h, C.byref(st), L.ptr(ga), L.ptr(gb_), L.ptr(gc_), B, S, L.SAMPLING_IDS[sampling_type],
                    ray_radius, L.ptr(save), L.ptr(g_density), L.ptr(g_color), L.ptr(g_penalty), L.ptr(post),
                    L.ptr(gpre), L.ptr(ghead_da), L.ptr(ghead_col), L.ptr(xes), L.ptr(xcol), L.stream_ptr(device)),
                    "field_backward")
            else:
                L.check(L.lib().neddf_field_backward_samples(
                    h, C.byref(st), L.ptr(ga), L.ptr(gb_), L.ptr(gc_), n, L.ptr(save), L.ptr(g_density),
                    L.ptr(g_color), L.ptr(g_penalty), L.ptr(post), L.ptr(gpre), L.ptr(ghead_da), L.ptr(ghead_col),
                    L.ptr(xes), L.ptr(xcol), L.stream_ptr(device)), "field_backward_samples")

        # weight gradients gW = X^T G over the 4N rows (linear.py:76-79) and bias gradients (sum over the value
        # rows): tensor-core split-K GEMMs of this library (csrc/wgrad.cu), written straight into the gradient
        # tensors - no library GEMM on the training path
        lib = L.lib()
        stream = L.stream_ptr(device)
        ws = getattr(net, "_wgrad_ws", None)
        if ws is None or ws.device != device:
            ws = torch.empty(int(lib.neddf_wgrad_workspace_bytes()) // 4, device=device, dtype=torch.float32)
            net._wgrad_ws = ws
        R = 4 * n

        def wgrad_into(out, row0, A, lda, ka, Bm, n_cols=256):
            """out[row0 : row0 + ka, :n_cols] = A[:, :ka]^T Bm, in 128-column tiles of A."""
            for c0 in range(0, ka, 128):
                kk = min(128, ka - c0)
                L.check(lib.neddf_wgrad(L.ptr(A), lda, c0, kk, L.ptr(Bm), 256, R,
                                        C.c_void_p(out.data_ptr() + 4 * (row0 + c0) * out.shape[1]), out.shape[1], n_cols,
                                        L.ptr(ws), stream), "wgrad")

        grads = []
        with torch.cuda.device(device):
            for l in range(n_hidden):
                if l == 0:
                    parts = [(xes, n_e0)]
                elif l < n_ddf:
                    parts = ([(xes, n_e0)] if (l - 1) in net.skips else []) + [(post[l - 1], 256)]
                elif l == n_ddf:
                    parts = [(xcol, off_h), (post[n_ddf - 1], 256)]
                else:
                    parts = [(post[l - 1], 256)]
                gW = torch.empty(sum(k for _, k in parts), 256, device=device, dtype=torch.float32)
                row0 = 0
                for X, k_in in parts:
                    wgrad_into(gW, row0, X, k_in, k_in, gpre[l])
                    row0 += k_in
                gb = torch.empty(256, device=device, dtype=torch.float32)
                L.check(lib.neddf_colsum_value_rows(L.ptr(gpre[l]), n, 4 * 256, L.ptr(gb), L.ptr(ws), stream), "colsum")
                grads += [gW, gb]
            # heads: gW^T [outs, 256] = ghead^T post (the 2- / 3-column head gradients are the A operand)
            gda_t = torch.empty(2, 256, device=device, dtype=torch.float32)
            wgrad_into(gda_t, 0, ghead_da, 2, 2, post[n_ddf - 1])
            gc_t = torch.empty(3, 256, device=device, dtype=torch.float32)
            wgrad_into(gc_t, 0, ghead_col, 4, 3, post[n_hidden - 1])
        b_da = ghead_da[:, 0, :].sum(0)
        grads += [gda_t[0].reshape(256, 1).contiguous(), b_da[0:1].contiguous(), gda_t[1].reshape(256, 1).contiguous(),
                  b_da[1:2].contiguous()]
        grads += [gc_t.t().contiguous(), ghead_col[:, 0, :3].sum(0)]
        return (None, None, None, None, None, None) + tuple(grads)


class EngineRangeError(FloatingPointError):
    """The tensor-core engine left fp16 range under engine "auto"; the network has switched itself to the fp32 engine
    and the caller (NeRFRender) re-runs the call.  Explicit engines ("tc", "tc2") raise plain FloatingPointError."""


class BaseNeuralField(nn.Module):
    """neddf/network/base_neuralfield.py:11-79."""

    @property
    def device(self) -> torch.device:
        return next(self.parameters()).device

    def set_iter(self, iter: int) -> None:
        pass

    def voxelize(self, field_name: str = "density", cube_range: float = 1.1, cube_resolution: int = 64,
                 chunk: int = 65536) -> np.ndarray:
        """Dense-grid evaluation (base_neuralfield.py:49-79); same point order as the reference."""
        with torch.set_grad_enabled(False):
            ids = np.linspace(-cube_range, cube_range, cube_resolution)
            zs, ys, xs = np.meshgrid(ids, ids, ids)
            pos = torch.from_numpy(np.stack([xs.reshape(-1), ys.reshape(-1), zs.reshape(-1)], 1).astype(np.float32))
            n = cube_resolution ** 3
            device = self.device
            result = np.zeros(n, np.float32)
            one_dir = torch.tensor([[1.0, 0.0, 0.0]])
            for i in range(0, n, chunk):
                j = min(n, i + chunk)
                p = pos[None, i:j, :].to(device)
                s = Sampling(p, one_dir.expand(j - i, -1)[None].to(device).contiguous(), torch.zeros_like(p))
                result[i:j] = self.forward(s)[field_name].view(-1).detach().cpu().numpy()
            return result.reshape(cube_resolution, cube_resolution, cube_resolution)


class NeDDF(BaseNeuralField):
    """Drop-in for neddf.network.NeDDF (neddf/network/neddf.py:21-326)."""

    def __init__(
        self,
        embed_pos_rank: int = 10,
        embed_dir_rank: int = 4,
        ddf_layer_count: int = 8,
        ddf_layer_width: int = 256,
        col_layer_count: int = 8,
        col_layer_width: int = 256,
        activation_type: str = "tanhExp",
        density_activation_type: str = "ReLU",
        d_near: float = 0.01,
        lowpass_alpha_offset: float = 10.0,
        skips: Optional[List[int]] = None,
        penalty_weight: Optional[Dict[str, float]] = None,
    ) -> None:
        super().__init__()
        input_ddf_dim = embed_pos_rank * 6
        input_col_dim = (embed_pos_rank + embed_dir_rank) * 6 + 3 + ddf_layer_width
        if skips is None:
            skips = [4]
        self.skips = [int(s) for s in skips]
        if activation_type not in L.ACT_IDS or density_activation_type not in L.ACT_IDS:
            raise KeyError(f"unknown activation {activation_type!r}/{density_activation_type!r}")  # neddf.py:107-118
        self.activation_type = activation_type
        self.density_activation_type = density_activation_type
        self.embed_pos_rank, self.embed_dir_rank = int(embed_pos_rank), int(embed_dir_rank)
        self.ddf_layer_count, self.col_layer_count = int(ddf_layer_count), int(col_layer_count)
        self.ddf_layer_width, self.col_layer_width = int(ddf_layer_width), int(col_layer_width)

        # identical construction order and shapes to neddf.py:129-145
        layers_ddf = [LinearGradLayer(input_ddf_dim, ddf_layer_width)]
        for layer_id in range(ddf_layer_count - 2):
            extra = input_ddf_dim if layer_id in self.skips else 0
            layers_ddf.append(LinearGradLayer(ddf_layer_width + extra, ddf_layer_width))
        layers_col = [LinearGradLayer(input_col_dim, col_layer_width)]
        for _ in range(col_layer_count - 2):
            layers_col.append(LinearGradLayer(col_layer_width, col_layer_width))
        self.layers_ddf = nn.ModuleList(layers_ddf)
        self.layers_col = nn.ModuleList(layers_col)
        self.layer_ddf_out = LinearGradLayer(ddf_layer_width, 1)
        self.layer_aux_out = LinearGradLayer(ddf_layer_width, 1)
        self.layer_col_out = LinearGradLayer(ddf_layer_width, 3)

        self.d_near = float(d_near)
        self.aux_grad_scale = 1.1
        self.distance_range_max = 2.0
        self.lowpass_alpha_offset = float(lowpass_alpha_offset)
        self.lowpass_alpha = float(lowpass_alpha_offset)
        if penalty_weight is None:
            penalty_weight = {"constraints_aux_grad": 0.05, "constraints_dDdt": 0.05, "constraints_color": 0.01,
                              "range_distance": 1.0, "range_aux_grad": 1.0}
        self.penalty_weight = {k: float(v) for k, v in dict(penalty_weight).items()}

        # kernel-side state
        self.engine = "auto"          # "auto" | "fp32" | "tc" | "tc2"
        self._range_fallback = False  # "auto" met an activation outside fp16 range: it resolves to fp32 from then on
        self._handle = None
        self._handle_device = None
        self._packed_key = None
        self._profile_events = None   # bench.py: list receiving (start, end, n_evaluations) CUDA events

    def resolved_engine(self, device=None) -> str:
        """Engine that will actually run for ``self.engine`` ("fp32" or "tc")."""
        h = self._field(torch.device(device) if device is not None else self.device)
        rc = L.check(L.lib().neddf_field_resolve_engine(h, self._engine_id()), "resolve_engine")
        return {1: "fp32", 2: "tc", 3: "tc2"}[rc]

    def _engine_id(self) -> int:
        """ABI engine id of the next launch: ``self.engine``, except that "auto" stays on the fp32 engine once the
        tensor-core engine has reported an activation outside fp16 range for this network (check_engine_status)."""
        return L.ENGINE_IDS["fp32" if (self._range_fallback and self.engine == "auto") else self.engine]

    # ------------------------------------------------------------------ kernel plumbing --
    def _ordered_layers(self) -> List[LinearGradLayer]:
        return list(self.layers_ddf) + list(self.layers_col) + [self.layer_ddf_out, self.layer_aux_out, self.layer_col_out]

    def _config_struct(self) -> L.FieldConfig:
        c = L.FieldConfig()
        c.embed_pos_rank, c.embed_dir_rank = self.embed_pos_rank, self.embed_dir_rank
        c.ddf_layer_count, c.ddf_layer_width = self.ddf_layer_count, self.ddf_layer_width
        c.col_layer_count, c.col_layer_width = self.col_layer_count, self.col_layer_width
        c.activation_type = L.ACT_IDS[self.activation_type]
        c.density_activation_type = L.ACT_IDS[self.density_activation_type]
        c.d_near = self.d_near
        if len(self.skips) > L.MAX_SKIPS:
            raise NotImplementedError("neddf_b200: more than 8 skip connections")
        c.n_skips = len(self.skips)
        for i, s in enumerate(self.skips):
            c.skips[i] = s
        for i, k in enumerate(L.PENALTY_KEYS):  # absent key -> unweighted, neddf.py:296-299
            c.penalty_weight[i] = self.penalty_weight.get(k, 1.0)
        return c

    def _state_struct(self) -> L.FieldState:
        """Per-call scalars: the warm-up schedule and the penalty weights, both read from the module on
        every forward like the reference does (neddf.py:296-299: absent key -> unweighted)."""
        pw = (C.c_float * L.N_PENALTY)(*[float(self.penalty_weight.get(k, 1.0)) for k in L.PENALTY_KEYS])
        return L.FieldState(float(self.aux_grad_scale), float(self.distance_range_max), float(self.lowpass_alpha), pw)

    def _release(self) -> None:
        if self._handle is not None:
            try:
                L.lib().neddf_field_destroy(self._handle)
            except Exception:  # interpreter shutdown
                pass
            self._handle = None
            self._packed_key = None

    def __del__(self):
        try:
            self._release()
        except Exception:  # interpreter shutdown: torch internals may already be gone
            pass

    def _field(self, device: torch.device):
        """Handle with weights packed for the parameters' current values."""
        lib = L.lib()
        if device.type != "cuda":
            raise RuntimeError("neddf_b200.NeDDF runs on CUDA devices only: move the module with .to('cuda') "
                               "(the hot path has no CPU implementation)")
        if self._handle is None or self._handle_device != device:
            self._release()
            h = C.c_void_p()
            with torch.cuda.device(device):
                cfg = self._config_struct()
                L.check(lib.neddf_field_create(C.byref(cfg), C.byref(h)), "field_create")
            self._handle, self._handle_device = h, device
        layers = self._ordered_layers()
        key = tuple((p.data_ptr(), p._version) for l in layers for p in (l.weight, l.bias))
        if key != self._packed_key:
            n = len(layers)
            ws = (C.c_void_p * n)(*[l.weight.data_ptr() for l in layers])
            bs = (C.c_void_p * n)(*[l.bias.data_ptr() for l in layers])
            for l in layers:
                if l.weight.dtype != torch.float32 or not l.weight.is_contiguous() or l.weight.device != device:
                    raise RuntimeError("neddf_b200: parameters must be contiguous fp32 tensors on the module's device")
            with torch.cuda.device(device):
                L.check(lib.neddf_field_set_weights(self._handle, ws, bs, n, L.stream_ptr(device)), "field_set_weights")
            self._packed_key = key
        return self._handle

    def _apply(self, fn, *a, **k):
        r = super()._apply(fn, *a, **k)
        self._packed_key = None  # .to()/.cuda() replaced the parameter storage
        return r

    def invalidate(self) -> None:
        """Force a re-pack of the kernel-layout weights on the next call.  Needed only after edits that
        bypass the parameters' version counters (``p.data.copy_(...)``, EMA swaps through ``.data``);
        optimiser steps, ``load_state_dict`` and ``.to()`` are detected automatically."""
        self._packed_key = None

    # the kernel handle is a process-local pointer: copies and pickles get a fresh one lazily
    def __getstate__(self):
        d = self.__dict__.copy()
        d["_handle"] = None
        d["_handle_device"] = None
        d["_packed_key"] = None
        d["_profile_events"] = None
        return d

    def __deepcopy__(self, memo):
        import copy
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__getstate__().items():
            new.__dict__[k] = copy.deepcopy(v, memo)
        return new

    # ------------------------------------------------------------------------- forward --
    def forward(self, sampling: Sampling) -> Dict[str, Tensor]:
        """NeDDF.forward (neddf.py:162-309): Sampling[B,S,3] -> distance, density, color,
        fields_penalty, aux_grad.  Under autograd: the differentiable fp32 path (gradients to the
        parameters through density / color / fields_penalty; distance and aux_grad are returned
        without a graph - no loss of the reference consumes them)."""
        pos = sampling.sample_pos
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            B, S = pos.shape[0], pos.shape[1]
            p3 = L.require_cuda_f32(pos.reshape(B, S, 3), "sample_pos")
            d3 = L.require_cuda_f32(sampling.sample_dir.reshape(B, S, 3), "sample_dir")
            v3 = L.require_cuda_f32(sampling.diag_variance.reshape(B, S, 3), "diag_variance")
            flat = [t for l in self._ordered_layers() for t in (l.weight, l.bias)]
            d, c, pnl, dist, aux = _FieldTrainFn.apply(self, p3, d3, v3, None, 0.0, *flat)
            return {"distance": dist, "density": d, "color": c, "fields_penalty": pnl, "aux_grad": aux}
        B, S = pos.shape[0], pos.shape[1]
        device = pos.device
        # reshape, not view: accept the expanded tensors the reference's point sampler returns
        p3 = L.require_cuda_f32(pos.reshape(-1, 3), "sample_pos")
        d3 = L.require_cuda_f32(sampling.sample_dir.reshape(-1, 3), "sample_dir")
        v3 = L.require_cuda_f32(sampling.diag_variance.reshape(-1, 3), "diag_variance")
        n = B * S
        out = {
            "distance": torch.empty(B, S, device=device, dtype=torch.float32),
            "density": torch.empty(B, S, device=device, dtype=torch.float32),
            "color": torch.empty(B, S, 3, device=device, dtype=torch.float32),
            "fields_penalty": torch.empty(B, S, device=device, dtype=torch.float32),
            "aux_grad": torch.empty(B, S, device=device, dtype=torch.float32),
        }
        h = self._field(device)
        st = self._state_struct()
        with torch.cuda.device(device):
            L.check(L.lib().neddf_field_forward(
                h, C.byref(st), L.ptr(p3), L.ptr(d3), L.ptr(v3), n, L.ptr(out["distance"]), L.ptr(out["density"]),
                L.ptr(out["color"]), L.ptr(out["fields_penalty"]), L.ptr(out["aux_grad"]), L.OUT_FULL,
                self._engine_id(), L.stream_ptr(device)), "field_forward")
        return out

    def forward_rays(self, ray_dir: Tensor, ray_orig: Tensor, dists: Tensor, sampling_type: str, ray_radius: float,
                     need_penalty: bool = True, need_aux: bool = True) -> Dict[str, Tensor]:
        """Same network with the sample geometry fused into the kernel prologue (no [N,3]
        Sampling tensors in HBM).  Used by NeRFRender.  Under autograd (training) it runs the
        differentiable fp32 path (_FieldTrainFn) and returns density / color / fields_penalty."""
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            ray_dir = L.require_cuda_f32(ray_dir, "ray_dir")
            ray_orig = L.require_cuda_f32(ray_orig, "ray_orig")
            dists = L.require_cuda_f32(dists, "dists")
            flat = [t for l in self._ordered_layers() for t in (l.weight, l.bias)]
            d, c, pnl = _FieldTrainFn.apply(self, ray_dir, ray_orig, dists, sampling_type, ray_radius, *flat)
            return {"density": d, "color": c, "fields_penalty": pnl}
        ray_dir = L.require_cuda_f32(ray_dir, "ray_dir")
        ray_orig = L.require_cuda_f32(ray_orig, "ray_orig")
        dists = L.require_cuda_f32(dists, "dists")
        B, S = dists.shape
        device = dists.device
        out = {
            "density": torch.empty(B, S, device=device, dtype=torch.float32),
            "color": torch.empty(B, S, 3, device=device, dtype=torch.float32),
        }
        if need_penalty:
            out["fields_penalty"] = torch.empty(B, S, device=device, dtype=torch.float32)
        if need_aux:
            out["distance"] = torch.empty(B, S, device=device, dtype=torch.float32)
            out["aux_grad"] = torch.empty(B, S, device=device, dtype=torch.float32)
        h = self._field(device)
        st = self._state_struct()
        prof = self._profile_events
        if prof is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(torch.cuda.current_stream(device))
        with torch.cuda.device(device):
            L.check(L.lib().neddf_field_forward_rays(
                h, C.byref(st), L.ptr(ray_dir), L.ptr(ray_orig), L.ptr(dists), B, S, L.SAMPLING_IDS[sampling_type],
                float(ray_radius), L.ptr(out.get("distance")), L.ptr(out["density"]), L.ptr(out["color"]),
                L.ptr(out.get("fields_penalty")), L.ptr(out.get("aux_grad")),
                L.OUT_FULL if need_penalty else L.OUT_EVAL, self._engine_id(), L.stream_ptr(device)),
                "field_forward_rays")
        if prof is not None:
            e1.record(torch.cuda.current_stream(device))
            prof.append((e0, e1, B * S))
        return out

    def forward_rays_segment(self, ray_dir: Tensor, ray_orig: Tensor, dists: Tensor, sampling_type: str, ray_radius: float,
                             edge0: int, seg_len: int, ray_index: Optional[Tensor], n_active: Optional[Tensor],
                             density: Tensor, color: Tensor) -> None:
        """One depth segment of the fine pass for early ray termination (neddf_field_forward_rays_segment):
        samples [edge0, edge0 + seg_len) of the rays listed in ``ray_index[:n_active]`` (device tensors, None =
        all rays); results are scattered into ``density`` [B,E] / ``color`` [B,E,3] in place.  No-grad only."""
        B, E = dists.shape
        device = dists.device
        h = self._field(device)
        st = self._state_struct()
        prof = self._profile_events
        if prof is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(torch.cuda.current_stream(device))
        with torch.cuda.device(device):
            L.check(L.lib().neddf_field_forward_rays_segment(
                h, C.byref(st), L.ptr(ray_dir), L.ptr(ray_orig), L.ptr(dists), B, E, L.SAMPLING_IDS[sampling_type],
                float(ray_radius), int(edge0), int(seg_len), L.ptr(ray_index), L.ptr(n_active), L.ptr(density),
                L.ptr(color), self._engine_id(), L.stream_ptr(device)), "field_forward_rays_segment")
        if prof is not None:
            e1.record(torch.cuda.current_stream(device))
            prof.append((e0, e1, None))  # the executed count lives on the device (NeRFRender.termination_stats)

    def check_engine_status(self) -> None:
        """Read-and-clear the engine's device status word (one sync).  Raises if the tensor-core
        engine met an activation outside fp16 range (its operands are fp16 hi+lo pairs)."""
        if self._handle is None:
            return
        v = C.c_int32(0)
        dev = self._handle_device
        with torch.cuda.device(dev):
            L.check(L.lib().neddf_field_status(self._handle, C.byref(v), L.stream_ptr(dev)), "field_status")
        if v.value & 4:
            if self.engine == "auto" and not self._range_fallback:
                self._range_fallback = True
                raise EngineRangeError(
                    "neddf_b200: tensor-core engine saw |activation| > 65504 (fp16 range of its split operands); "
                    "engine 'auto' now resolves to the fp32 engine for this network and the call is re-run")
            raise FloatingPointError(
                "neddf_b200: tensor-core engine saw |activation| > 65504 (fp16 range of its split operands); "
                "results of the last calls are invalid - use set_engine('fp32') or 'auto' for this network")

    def set_iter(self, iter: int) -> None:
        """Warm-up schedule (neddf.py:311-326); -1 = evaluation."""
        if iter == -1:
            self.aux_grad_scale = 1.1
            self.distance_range_max = 2.0
            self.lowpass_alpha = float(self.embed_pos_rank)
        else:
            self.aux_grad_scale = min(1.1, max(0.01, 0.0001 * iter))
            self.distance_range_max = min(2.0, 2.0 + 0.0001 * iter)
            self.lowpass_alpha = self.lowpass_alpha_offset + 0.001 * iter
