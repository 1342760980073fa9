"""CPU: the hand-derived field backward (tests/manual_backward.py, the CPU twin of the CUDA backward
kernel) against torch autograd through the oracle, in fp64 (exact derivation check) and fp32."""
import pytest
import torch

from oracle import neddf_oracle as orc
from tests.helpers import Case, nerr
from tests.manual_backward import field_backward


def _setup(name, dtype, n_rays=6, seed=0):
    c = Case(name)
    d, o = orc.make_rays(c.t("uv")[:n_rays], c.cam, dtype)
    pos, dd, var = orc.make_samples(c.rc, d, o, c.t("dists_fine")[:n_rays, ::6].to(dtype))
    g = torch.Generator().manual_seed(seed)
    B, S = pos.shape[:2]
    gd = torch.randn(B, S, generator=g, dtype=dtype)
    gc = torch.randn(B, S, 3, generator=g, dtype=dtype)
    gp = torch.randn(B, S, generator=g, dtype=dtype)
    P = {k: v.to(dtype) for k, v in c.p_fine.items()}
    return c, P, pos, dd.contiguous(), var, gd, gc, gp


@pytest.mark.parametrize("name", ["train", "bunny", "point", "leaky"])
def test_manual_backward_matches_autograd_fp64(name):
    c, P, pos, dd, var, gd, gc, gp = _setup(name, torch.float64)
    Pg = {k: v.clone().requires_grad_(True) for k, v in P.items()}
    out = orc.field_forward(Pg, c.fc, c.st, pos, dd, var)
    loss = (out["density"] * gd).sum() + (out["color"] * gc).sum() + (out["fields_penalty"] * gp).sum()
    loss.backward()
    grads = field_backward(P, c.fc, c.st, pos, dd, var, gd, gc, gp)
    assert set(grads) == set(P)
    for k, v in Pg.items():
        assert nerr(grads[k].numpy(), v.grad.numpy()) < 1e-7, k


def test_manual_backward_fp32_noise_floor():
    c, P, pos, dd, var, gd, gc, gp = _setup("train", torch.float32)
    Pg = {k: v.clone().requires_grad_(True) for k, v in P.items()}
    out = orc.field_forward(Pg, c.fc, c.st, pos, dd, var)
    ((out["density"] * gd).sum() + (out["color"] * gc).sum() + (out["fields_penalty"] * gp).sum()).backward()
    grads = field_backward(P, c.fc, c.st, pos, dd, var, gd, gc, gp)
    for k, v in Pg.items():
        assert nerr(grads[k].numpy(), v.grad.numpy()) < 5e-4, k
