"""CPU, world_size 2 (gloo): the sharding + single all-gather logic of neddf_b200.dist."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from neddf_b200.dist import gather_tiles, render_image_sharded, shard_range, tile_capacity


def test_shard_range_tiles_exactly():
    for n in (0, 1, 7, 640000, 640001, 250000):
        for world in (1, 2, 3, 4, 8):
            pos = 0
            for r in range(world):
                first, count = shard_range(n, world, r)
                assert first == pos and 0 <= count <= tile_capacity(n, world)
                pos += count
            assert pos == n


class _FakeRender:
    """Stands in for NeRFRender.render_pixels: pixel p -> (p, 2p, 3p | p/7)."""

    def render_pixels(self, width, height, camera, target_types, downsampling, first, count, uniforms=None):
        p = torch.arange(first, first + count, dtype=torch.float32)
        out = {"color": torch.stack([p, 2 * p, 3 * p], 1), "depth": (p / 7).reshape(-1, 1)}
        if uniforms is not None:
            out["color"] = out["color"] + uniforms[0][:, :1]
        return {k: out[k] for k in target_types}


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, w, h):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = w * h
        first, count = shard_range(n, world, rank)
        tile = torch.arange(first, first + count, dtype=torch.float32).reshape(-1, 1).repeat(1, 4)
        full = gather_tiles(tile, n)
        assert full.shape == (n, 4)
        assert torch.equal(full[:, 0], torch.arange(n, dtype=torch.float32))
        u = (torch.arange(n, dtype=torch.float32).reshape(-1, 1) * 0.5, torch.zeros(n, 1))
        img = render_image_sharded(_FakeRender(), w, h, None, ["color", "depth"], 1, uniforms=u)
        p = torch.arange(n, dtype=torch.float32)
        assert img["color"].shape == (h, w, 3) and img["depth"].shape == (h, w, 1)
        assert torch.equal(img["color"].reshape(n, 3), torch.stack([p, 2 * p, 3 * p], 1) + 0.5 * p.reshape(-1, 1))
        assert torch.equal(img["depth"].reshape(n), p / 7)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("shape", [(8, 6), (7, 5)])  # even and ragged split
def test_sharded_render_two_ranks(shape):
    mp.spawn(_worker, args=(2, _free_port(), shape[0], shape[1]), nprocs=2, join=True)
