"""Ray / Sampling containers with the reference's attribute names.

Reference: neddf/ray/ray.py:8-86 (Ray), neddf/ray/sampling.py:5-42 (Sampling).  The geometry
itself (get_sampling_points / get_sampling_cones, ray.py:88-194) runs in CUDA
(neddf_make_samples, or fused into the field kernel's prologue).
"""
from typing import Tuple

import torch
from torch import Tensor

from . import _lib as L

CONE_RAY_RADIUS = 1.0 / 1111 / (12.0 ** 0.5)  # neddf/render/nerf_render.py:145


class Sampling:
    """[batch, samples, 3] sample positions, directions and diagonal variances."""

    def __init__(self, sample_pos: Tensor, sample_dir: Tensor, diag_variance: Tensor) -> None:
        assert sample_pos.shape == sample_dir.shape  # sampling.py:33-34
        assert sample_pos.shape == diag_variance.shape
        self.sample_pos = sample_pos
        self.sample_dir = sample_dir
        self.diag_variance = diag_variance

    @property
    def device(self) -> torch.device:
        return self.sample_pos.device


class Ray:
    def __init__(self, ray_dir: Tensor, ray_orig: Tensor, uv: Tensor) -> None:
        self.single = ray_dir.dim() == 1
        assert ray_orig.shape == ray_dir.shape  # ray.py:44-48
        if self.single:
            assert uv.shape == (2,)
        else:
            assert uv.shape == (ray_orig.shape[0], 2)
        self.ray_dir, self.ray_orig, self.uv = ray_dir, ray_orig, uv

    @property
    def device(self) -> torch.device:
        return self.ray_dir.device

    def __len__(self) -> int:
        return 1 if self.single else self.ray_dir.shape[0]

    def __getitem__(self, item: int) -> Tuple[Tensor, Tensor]:
        if self.single:
            return (self.ray_dir, self.ray_orig)
        return (self.ray_dir[item, :], self.ray_orig[item, :])

    def _samples(self, dists: Tensor, sampling_type: str, ray_radius: float) -> Sampling:
        batch, count = dists.shape
        assert batch == self.ray_dir.shape[0]  # ray.py:111,152
        d = L.require_cuda_f32(self.ray_dir, "ray_dir")
        o = L.require_cuda_f32(self.ray_orig, "ray_orig")
        t = L.require_cuda_f32(dists, "dists")
        pos = torch.empty(batch, count, 3, device=t.device, dtype=torch.float32)
        sdir = torch.empty_like(pos)
        var = torch.empty_like(pos)
        with torch.cuda.device(t.device):
            L.check(L.lib().neddf_make_samples(L.ptr(d), L.ptr(o), L.ptr(t), batch, count,
                                               L.SAMPLING_IDS[sampling_type], float(ray_radius), L.ptr(pos),
                                               L.ptr(sdir), L.ptr(var), L.stream_ptr(t.device)), "make_samples")
        return Sampling(pos, sdir, var)

    def get_sampling_points(self, dists: Tensor) -> Sampling:
        """pos = o + d t, zero variance (ray.py:88-126)."""
        return self._samples(dists, "point", 0.0)

    def get_sampling_cones(self, dists: Tensor, ray_radius: float = 1e-3) -> Sampling:
        """Conical-frustum mean and diagonal variance per edge (ray.py:128-194)."""
        return self._samples(dists, "cone", ray_radius)
