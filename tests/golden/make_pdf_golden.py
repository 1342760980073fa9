"""Golden vectors of BaseNeuralRender.sample_pdf(cat_coarse=False) from the REAL reference
(neddf/render/base_neural_render.py:27-115), imported from /root/reference in the build container with the
hydra / omegaconf stand-ins of tests/golden/_refstub; the internal torch.rand is replaced by recorded uniforms.

    python tests/golden/make_pdf_golden.py      ->  tests/golden/case_pdf_nocat.npz
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "_refstub"))
sys.path.insert(0, "/root/reference")

from neddf.render.base_neural_render import BaseNeuralRender  # noqa: E402


class _R(BaseNeuralRender):  # the abstract bits are not on this path
    def render_rays(self, *a, **k):
        raise NotImplementedError

    def get_parameters_list(self):
        return []

    def render_image(self, *a, **k):
        raise NotImplementedError

    def render_field_slice(self, *a, **k):
        raise NotImplementedError

    def get_network(self):
        return None


g = torch.Generator().manual_seed(11)
B, E, F = 7, 33, 20
dists = torch.sort(torch.rand(B, E, generator=g) * 4 + 2, dim=1).values
w = torch.rand(B, E - 1, generator=g) ** 3
w[2, 5] = -0.2           # sanitised in place (:52-55)
w[4, 9] = float("nan")
u = torch.rand(B, F, generator=g)
orig = torch.rand
torch.rand = lambda *s, **k: u.clone()
try:
    r = _R()
    w_in = w.clone()
    out = r.sample_pdf(dists.clone(), w_in, F, cat_coarse=False)
finally:
    torch.rand = orig
np.savez(os.path.join(HERE, "case_pdf_nocat.npz"), dists=dists.numpy(), weights=w.numpy(), weights_after=w_in.numpy(),
         u=u.numpy(), out=out.numpy())
print("wrote case_pdf_nocat.npz", out.shape)
