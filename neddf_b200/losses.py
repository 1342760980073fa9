"""The reference's training objective as one CUDA launch.

Reference: neddf/loss/base_loss.py:45-85 (weight / weight_coarse wrapper and the ``*_coarse`` keys),
color_loss.py:41-55, mask_bce_loss.py:41-59, fields_constraint_loss.py:40-54, summed by the trainer
(nerf_trainer.py:118-121).  ``neddf_render_loss`` evaluates the six weighted terms and, in the same pass, the
gradients of their sum w.r.t. the render outputs, so the backward of the loss is a scalar multiply per output
(no [B]-sized autograd graph of ~30 element-wise kernels).

``ColorLoss`` / ``MaskBCELoss`` / ``FieldsConstraintLoss`` keep the reference's constructor and call
signature (``install(patch_trainer=True)`` binds them over ``neddf.loss``); ``RenderLoss`` computes all of
them with a single launch.
"""
from typing import Dict, Optional, Tuple

import torch
from torch import Tensor, nn

from . import _lib as L

_KEYS = ("color", "color_coarse", "mask", "mask_coarse", "fields_penalty", "fields_penalty_coarse")


class _RenderLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, weights: Tensor, t_color, t_mask, color, color_c, trans, trans_c, pen, pen_c):
        outs = [color, color_c, trans, trans_c, pen, pen_c]
        ref = next(t for t in outs if t is not None)
        device, B = ref.device, ref.shape[0]

        def f32(t):
            return None if t is None else L.require_cuda_f32(t.detach(), "loss input")

        ins = [f32(t) for t in outs]
        tc, tm = f32(t_color), f32(t_mask)
        terms = torch.zeros(6, device=device, dtype=torch.float32)
        grads = [None if t is None else torch.empty_like(t) for t in ins]
        with torch.cuda.device(device):
            L.check(L.lib().neddf_render_loss(L.ptr(ins[0]), L.ptr(ins[1]), L.ptr(ins[2]), L.ptr(ins[3]), L.ptr(ins[4]),
                                              L.ptr(ins[5]), L.ptr(tc), L.ptr(tm), B, L.ptr(weights), L.ptr(terms),
                                              L.ptr(grads[0]), L.ptr(grads[1]), L.ptr(grads[2]), L.ptr(grads[3]),
                                              L.ptr(grads[4]), L.ptr(grads[5]), L.stream_ptr(device)), "render_loss")
        ctx.grads = grads
        ctx.shapes = [None if t is None else t.shape for t in outs]
        return terms

    @staticmethod
    def backward(ctx, g_terms):
        res = []
        order = (0, 1, 2, 3, 4, 5)  # output k is fed by term k only
        for k, g in zip(order, ctx.grads):
            res.append(None if g is None else (g * g_terms[k]).reshape(ctx.shapes[k]))
        return (None, None, None) + tuple(res)


class RenderLoss(nn.Module):
    """All loss terms of config/loss/neddf_loss.yaml in one launch.  Returns the reference's ``loss_dict``
    (keys color, color_coarse, mask, mask_coarse, fields_penalty, fields_penalty_coarse; a term whose weight is
    0 is left out like base_loss.py:76 does for weight_coarse == 0)."""

    def __init__(self, color: Tuple[float, float] = (1.0, 0.1), mask: Tuple[float, float] = (0.05, 0.005),
                 fields_penalty: Tuple[float, float] = (0.01, 0.01)) -> None:
        super().__init__()
        self.weights = [float(color[0]), float(color[1]), float(mask[0]), float(mask[1]),
                        float(fields_penalty[0]), float(fields_penalty[1])]
        self._w_dev: Optional[Tensor] = None

    def _w(self, device) -> Tensor:
        if self._w_dev is None or self._w_dev.device != device:
            self._w_dev = torch.tensor(self.weights, device=device, dtype=torch.float32)
        return self._w_dev

    def forward(self, outputs: Dict[str, Tensor], targets: Dict[str, Tensor]) -> Dict[str, Tensor]:
        w = self.weights
        need = lambda k: w[k] != 0.0  # noqa: E731
        get = lambda key, k: outputs[key] if need(k) else None  # noqa: E731  (KeyError like the reference's assert)
        color, color_c = get("color", 0), get("color_coarse", 1)
        trans, trans_c = get("transmittance", 2), get("transmittance_coarse", 3)
        pen, pen_c = get("fields_penalty", 4), get("fields_penalty_coarse", 5)
        t_color = targets["color"] if (need(0) or need(1)) else None
        t_mask = targets["mask"] if (need(2) or need(3)) else None
        if t_mask is not None:
            t_mask = t_mask.reshape(-1)
        ref = next(t for t in (color, color_c, trans, trans_c, pen, pen_c) if t is not None)
        terms = _RenderLossFn.apply(self._w(ref.device), t_color, t_mask, color, color_c, trans, trans_c, pen, pen_c)
        return {k: terms[i] for i, k in enumerate(_KEYS) if need(i)}


class _SingleTerm(nn.Module):
    """One of the reference's loss classes on top of the fused kernel (its own two terms only)."""
    _slot = 0

    def __init__(self, weight: float = 1.0, weight_coarse: float = 0.1) -> None:
        super().__init__()
        self.weight, self.weight_coarse = float(weight), float(weight_coarse)
        args = {"color": (0.0, 0.0), "mask": (0.0, 0.0), "fields_penalty": (0.0, 0.0)}
        args[("color", "mask", "fields_penalty")[self._slot]] = (self.weight, self.weight_coarse)
        self._fused = RenderLoss(**args)

    def forward(self, outputs: Dict[str, Tensor], targets: Dict[str, Tensor]) -> Dict[str, Tensor]:
        d = self._fused(outputs, targets)
        if self.weight == 0.0:  # base_loss.py:70-72 always returns the fine key, even with weight 0
            key = _KEYS[2 * self._slot]
            d[key] = torch.zeros((), device=next(iter(outputs.values())).device)
        return d


class ColorLoss(_SingleTerm):
    """neddf.loss.ColorLoss (color_loss.py:10-55)."""
    _slot = 0


class MaskBCELoss(_SingleTerm):
    """neddf.loss.MaskBCELoss (mask_bce_loss.py:10-59)."""
    _slot = 1


class FieldsConstraintLoss(_SingleTerm):
    """neddf.loss.FieldsConstraintLoss (fields_constraint_loss.py:10-54)."""
    _slot = 2
