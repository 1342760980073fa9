"""Hand-derived backward of NeDDF.forward, written the way the CUDA backward kernel computes it
(layer by layer in reverse, explicit head / penalty derivatives).  Test infrastructure: it pins
the derivation against torch autograd through the oracle (tests/test_manual_backward.py) so that the
CUDA port has a line-by-line CPU twin.

Upstream gradients: g_density[N], g_color[N,3], g_penalty[N] (what the compositing backward
produces); `distance` and `aux_grad` are not consumed by render_rays' losses.
Returns gradients of every parameter (same keys as the oracle's parameter dict).
"""
from typing import Dict

import torch
from torch import Tensor

from oracle import neddf_oracle as orc


def _act_derivs(name: str, x: Tensor):
    """f(x), f'(x), f''(x) with the reference's masks (tanh_exp.py:38-53, relu.py, leaky_relu.py)."""
    if name == "tanhExp":
        big = x > 20.0
        ex = torch.exp(x)
        tx = torch.tanh(ex)
        y = torch.where(big, x, x * tx)
        d1 = torch.where(big, torch.ones_like(x), tx - x * ex * (tx * tx - 1))
        d2 = torch.where(big, torch.zeros_like(x), ex * (-x + 2 * ex * x * tx - 2) * (tx * tx - 1))
        return y, d1, d2
    if name == "ReLU":
        s = (x >= 0).to(x.dtype)
        return x * s, s, torch.zeros_like(x)
    if name == "LeakyReLU":
        s = torch.where(x < 0, torch.full_like(x, 0.01), torch.ones_like(x))
        return x * s, s, torch.zeros_like(x)
    raise ValueError(name)


def _density_act_deriv(name: str, z: Tensor) -> Tensor:
    if name == "ReLU":
        return (z > 0).to(z.dtype)
    if name == "LeakyReLU":
        return torch.where(z > 0, torch.ones_like(z), torch.full_like(z, 0.01))
    ex = torch.exp(z)  # tanhExp, nn_module/tanh_exp.py
    tx = torch.tanh(ex)
    return torch.where(z > 20.0, torch.ones_like(z), tx - z * ex * (tx * tx - 1))


def field_backward(P: Dict[str, Tensor], cfg: orc.FieldConfig, st: orc.FieldState, pos: Tensor, dirs: Tensor,
                   var: Tensor, g_density: Tensor, g_color: Tensor, g_penalty: Tensor) -> Dict[str, Tensor]:
    n = pos.shape[0] * pos.shape[1]
    dt = pos.dtype
    x3, d3, v3 = pos.reshape(n, 3), dirs.reshape(n, 3), var.reshape(n, 3)
    E = cfg.embed_pos_rank
    act = cfg.activation_type
    pw = [cfg.penalty_weight.get(k, 1.0) for k in ("constraints_aux_grad", "constraints_dDdt", "range_distance",
                                                   "range_aux_grad", "range_color", "constraints_color")]
    gsig, gcol, gpen = g_density.reshape(n, 1), g_color.reshape(n, 3), g_penalty.reshape(n, 1)

    # ---------------- forward with stored pre-activations (what the training forward keeps) ----------
    s_grad = (2.0 / (2.0 ** torch.arange(E, dtype=dt))).reshape(E, 1).expand(E, 3).reshape(1, 3 * E)
    s_low = orc.lowpass_scale(E, st.lowpass_alpha, dt).reshape(E, 1).expand(E, 3).reshape(1, 3 * E)
    w_pe = orc.pe_weights(v3, E)
    es, Jes = orc.pe_jacobian(x3, s_grad * s_low * w_pe, E)
    e0, Je0 = orc.pe_jacobian(x3, s_low * w_pe, E)
    ed = orc.pe_plain(d3, cfg.embed_dir_rank)
    n_ddf, n_col = cfg.ddf_layer_count - 1, cfg.col_layer_count - 1
    ins, pres = [], []  # per hidden layer: input (x[N,K], J[N,3,K]) and pre-activation (x, G)
    h, hJ = es, Jes
    for l in range(n_ddf):
        ins.append((h, hJ))
        x = h @ P[f"layers_ddf.{l}.weight"] + P[f"layers_ddf.{l}.bias"]
        G = hJ @ P[f"layers_ddf.{l}.weight"]
        pres.append((x, G))
        y, d1, _ = _act_derivs(act, x)
        h, hJ = y, d1.unsqueeze(1) * G
        if l in cfg.skips:
            h, hJ = torch.cat([es, h], 1), torch.cat([Jes, hJ], 2)
    feat, featJ = h, hJ
    wd, wa = P["layer_ddf_out.weight"], P["layer_aux_out.weight"]
    ddf_out = feat @ wd + P["layer_ddf_out.bias"]
    ddf_J = (featJ @ wd)[:, :, 0]
    aux_out = feat @ wa + P["layer_aux_out.bias"]
    aux_J = (featJ @ wa)[:, :, 0]
    big = ddf_out > 20.0
    sp = torch.where(big, ddf_out, torch.log(1 + torch.exp(ddf_out)))
    sp1 = torch.where(big, torch.ones_like(ddf_out), 1 / (1 + torch.exp(-ddf_out)))
    sp2 = torch.where(big, torch.zeros_like(ddf_out), (1 - sp1) * sp1)
    distance = sp + cfg.d_near
    grad_d = sp1 * ddf_J
    t = (1 + torch.tanh(aux_out * 0.5)) * 0.5
    t1 = t * (1 - t)
    t2 = t1 * (1 - 2 * t)
    s = st.aux_grad_scale
    aux = s * t
    aux_gg = s * t1 * aux_J
    n2 = (grad_d * grad_d).sum(1, keepdim=True)
    grad_norm = torch.sqrt(n2)
    dDdt = torch.sqrt(n2 + aux * aux)
    dist_inv = 1 / distance
    z = dist_inv * (1 - dDdt)
    q = 1 / (grad_norm + 1e-7)
    normal = q * grad_d
    c, cJ = torch.cat([e0, ed, normal, feat], 1), torch.cat([Je0, torch.zeros(n, 3, ed.shape[1] + 3, dtype=dt), featJ], 2)
    off_h = c.shape[1] - feat.shape[1]
    for l in range(n_col):
        ins.append((c, cJ))
        x = c @ P[f"layers_col.{l}.weight"] + P[f"layers_col.{l}.bias"]
        G = cJ @ P[f"layers_col.{l}.weight"]
        pres.append((x, G))
        y, d1, _ = _act_derivs(act, x)
        c, cJ = y, d1.unsqueeze(1) * G
    wc = P["layer_col_out.weight"]
    color = c @ wc + P["layer_col_out.bias"]
    colorJ = cJ @ wc  # [N,3,3]

    grads: Dict[str, Tensor] = {}

    # ---------------- colour head + colour penalties ---------------------------------------------------
    rc = torch.relu(-color) + torch.relu(color - 1)
    g_colv = gcol + gpen * pw[4] * 2 * rc * ((color > 1).to(dt) - (color < 0).to(dt))
    dot = (colorJ * grad_d.unsqueeze(2)).sum(1)  # [N,3]
    g_colJ = (gpen * pw[5] * 2 * dot).unsqueeze(1) * grad_d.unsqueeze(2)  # [N,3(i),3(c)]
    grads["layer_col_out.weight"] = c.T @ g_colv + torch.einsum("nik,nic->kc", cJ, g_colJ)
    grads["layer_col_out.bias"] = g_colv.sum(0)
    gy, gG = g_colv @ wc.T, g_colJ @ wc.T  # gradient wrt the colour trunk's last post-activation

    # ---------------- colour layers in reverse ------------------------------------------------------------
    def layer_backward(idx, name, gy, gG):
        x, G = pres[idx]
        xin, Jin = ins[idx]
        _, d1, d2 = _act_derivs(act, x)
        gx = gy * d1 + (gG * G).sum(1) * d2  # tanh_exp.py:84
        gGp = gG * d1.unsqueeze(1)           # tanh_exp.py:85
        W = P[name + ".weight"]
        grads[name + ".weight"] = xin.T @ gx + torch.einsum("nik,nic->kc", Jin, gGp)  # linear.py:76-79
        grads[name + ".bias"] = gx.sum(0)
        return gx @ W.T, gGp @ W.T  # linear.py:72-75

    for l in reversed(range(n_col)):
        gy, gG = layer_backward(n_ddf + l, f"layers_col.{l}", gy, gG)
    # colour input = [E0 | D | normal.detach() | feat]: only the feature part carries gradient
    g_feat, g_featJ = gy[:, off_h:], gG[:, :, off_h:]

    # ---------------- distance / aux heads, density, penalties ------------------------------------------
    g_z = gsig * _density_act_deriv(cfg.density_activation_type, z)
    g_dist_inv = g_z * (1 - dDdt)
    g_dDdt = -g_z * dist_inv + gpen * pw[1] * 2 * torch.relu(dDdt - 1)
    g_distance = -g_dist_inv * dist_inv * dist_inv
    safe = lambda v: torch.where(v > 0, v, torch.ones_like(v))
    g_n2 = torch.where(dDdt > 0, g_dDdt / (2 * safe(dDdt)), torch.zeros_like(dDdt))
    g_aux = torch.where(dDdt > 0, g_dDdt * aux / safe(dDdt), torch.zeros_like(dDdt))
    d2v = (aux_gg * normal).sum(1, keepdim=True)
    rest = 3 * aux * dist_inv
    A = aux * grad_norm * distance
    g_d2 = gpen * pw[0] * A * 2 * (d2v - rest)
    g_aux = g_aux + (-g_d2) * 3 * dist_inv
    g_aux_gg = g_d2 * normal
    g_normal = g_d2 * aux_gg
    g_grad_d = g_normal * q
    g_grad_norm = -(g_normal * grad_d).sum(1, keepdim=True) * q * q
    g_n2 = g_n2 + torch.where(grad_norm > 0, g_grad_norm / (2 * safe(grad_norm)), torch.zeros_like(grad_norm))
    g_grad_d = g_grad_d + g_n2 * 2 * grad_d
    g_t = g_aux * s
    g_t1 = (g_aux_gg * s * aux_J).sum(1, keepdim=True)
    g_aux_J = g_aux_gg * s * t1
    ra = torch.relu(-4.6 - aux_out) + torch.relu(aux_out - 4.6)
    g_aux_out = g_t * t1 + g_t1 * t2 + gpen * pw[3] * 2 * ra * ((aux_out > 4.6).to(dt) - (aux_out < -4.6).to(dt))
    g_sp1 = (g_grad_d * ddf_J).sum(1, keepdim=True)
    g_ddf_J = g_grad_d * sp1
    rd = torch.relu(-4.6 - ddf_out) + torch.relu(ddf_out - st.distance_range_max)
    g_ddf_out = g_distance * sp1 + g_sp1 * sp2 + gpen * pw[2] * 2 * rd * (
        (ddf_out > st.distance_range_max).to(dt) - (ddf_out < -4.6).to(dt))
    grads["layer_ddf_out.weight"] = feat.T @ g_ddf_out + torch.einsum("nik,ni->k", featJ, g_ddf_J).unsqueeze(1)
    grads["layer_ddf_out.bias"] = g_ddf_out.sum(0)
    grads["layer_aux_out.weight"] = feat.T @ g_aux_out + torch.einsum("nik,ni->k", featJ, g_aux_J).unsqueeze(1)
    grads["layer_aux_out.bias"] = g_aux_out.sum(0)
    gy = g_feat + g_ddf_out @ wd.T + g_aux_out @ wa.T
    gG = g_featJ + g_ddf_J.unsqueeze(2) * wd.T.unsqueeze(0) + g_aux_J.unsqueeze(2) * wa.T.unsqueeze(0)

    # ---------------- distance trunk in reverse ---------------------------------------------------------------
    n_es = es.shape[1]
    for l in reversed(range(n_ddf)):
        if l in cfg.skips:  # this layer's output was concatenated behind E_s: drop the E_s part
            gy, gG = gy[:, n_es:], gG[:, :, n_es:]
        gy, gG = layer_backward(l, f"layers_ddf.{l}", gy, gG)
    return grads
