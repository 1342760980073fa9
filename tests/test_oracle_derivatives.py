"""CPU: the oracle's forward-mode Jacobians against central finite differences in fp64 - the
reference's own strategy for its with_grad modules (tests/nn_module/with_grad/*.py),
applied to the whole restated field instead of one layer at a time."""
import pytest
import torch

from oracle import neddf_oracle as orc


def _setup(act, n=6, seed=0):
    cfg = orc.FieldConfig(activation_type=act)
    P = {k: v.double() for k, v in orc.init_params(cfg, seed=seed, bias_std=0.1).items()}
    st = orc.FieldState()
    g = torch.Generator().manual_seed(seed + 1)
    pos = (torch.rand(1, n, 3, generator=g, dtype=torch.float64) - 0.5) * 1.6
    dirs = torch.nn.functional.normalize(torch.randn(1, n, 3, generator=g, dtype=torch.float64), dim=-1)
    var = torch.rand(1, n, 3, generator=g, dtype=torch.float64) * 1e-3
    return cfg, P, st, pos, dirs, var


def _taps(cfg, P, st, pos, dirs, var):
    t = {}
    orc.field_forward(P, cfg, st, pos, dirs, var, taps=t)
    return t


@pytest.mark.parametrize("act", ["tanhExp", "LeakyReLU"])
def test_jacobians_are_derivatives_wrt_position(act):
    cfg, P, st, pos, dirs, var = _setup(act)
    base = _taps(cfg, P, st, pos, dirs, var)
    h = 1e-6
    # (value tap, Jacobian tap): Jacobian layout [N, 3, C] = d value[N, C] / d pos[N, i]
    pairs = [("embed_pos_scaled", "embed_pos_scaled_J"), ("embed_pos", "embed_pos_J"), ("ddf0_x", "ddf0_J"),
             ("ddf_out", "ddf_outJ"), ("aux_out", "aux_outJ")]
    last_ddf = max(int(k[3:-2]) for k in base if k.startswith("ddf") and k.endswith("_x") and k[3:-2].isdigit())
    pairs.append((f"ddf{last_ddf}_x", f"ddf{last_ddf}_J"))
    for i in range(3):
        dp = torch.zeros_like(pos)
        dp[..., i] = h
        hi, lo = _taps(cfg, P, st, pos + dp, dirs, var), _taps(cfg, P, st, pos - dp, dirs, var)
        for vk, jk in pairs:
            fd = (hi[vk] - lo[vk]) / (2 * h)
            J = base[jk]
            Ji = J[:, i] if J.dim() == 3 else J[:, i:i + 1]
            fd = fd.reshape(Ji.shape)
            scale = max(float(Ji.abs().max()), 1e-12)
            assert float((fd - Ji).abs().max()) / scale < 1e-6, (act, vk, i)


def test_density_follows_the_distance_gradient():
    """density = act((1 - |[grad D, aux]|) / distance) (neddf.py:234-240) with grad D the true
    spatial derivative of the returned distance."""
    cfg, P, st, pos, dirs, var = _setup("tanhExp", n=5, seed=3)
    out = orc.field_forward(P, cfg, st, pos, dirs, var)
    h = 1e-6
    grad = []
    for i in range(3):
        dp = torch.zeros_like(pos)
        dp[..., i] = h
        grad.append((orc.field_forward(P, cfg, st, pos + dp, dirs, var)["distance"]
                     - orc.field_forward(P, cfg, st, pos - dp, dirs, var)["distance"]) / (2 * h))
    grad = torch.stack(grad, -1)
    norm = torch.sqrt((grad * grad).sum(-1) + out["aux_grad"] ** 2)
    z = (1 - norm) / out["distance"]
    expect = torch.relu(z) if cfg.density_activation_type == "ReLU" else z
    if cfg.density_activation_type == "ReLU":
        assert float((out["density"] - expect).abs().max()) < 1e-6 * max(1.0, float(expect.abs().max()))
