#!/usr/bin/env python
"""Golden chunk of the BENCHMARKED workload from the REAL reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_bench_golden.py

bench.py's network / render configuration (BASELINE.json configs[1]: NeDDF 8x256 + 4x256 tanhExp, ReLU density, cone
sampling, 64 + 128 samples), its seeded random-init weights (bench.seeded_params, weights (ii) of SURVEY 8(d)) and its
synthetic camera 0; 1024 random pixels of the 800 x 800 frame through the reference's own NeRFRender.render_rays with
recorded uniforms.  Stored: inputs, the ten integrated outputs, the fine edge distances -> case_bench.npz.  The
weights are not stored: bench.seeded_params() regenerates them (tests/test_bench_contract.py pins that stream)."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402  (sets sys.path for the reference + stubs + repo)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    render = mg.build_render(bench.NET_CFG, bench.RENDER_CFG)
    sd, _ = bench.seeded_state_dict()
    print("load:", render.load_state_dict(sd))
    cam = mg.synthetic_camera(0, w=bench.W, h=bench.H)
    R, T, calib = bench.synthetic_pose(0)
    assert np.abs(cam.R.detach().numpy() - R).max() < 1e-6 and np.abs(cam.T.detach().numpy() - T).max() < 1e-6
    g = torch.Generator().manual_seed(1024)
    uv = torch.stack([torch.randint(0, bench.W, (1024,), generator=g), torch.randint(0, bench.H, (1024,), generator=g)], 1)
    res = mg.run_case(render, cam, uv, seed=1024, it=-1)
    out = {k: v for k, v in res.items() if not k.startswith("field_")}
    out["cfg"] = json.dumps({"network": bench.NET_CFG, "render": bench.RENDER_CFG, "weights": "bench.seeded_params()",
                             "pose": "bench.synthetic_pose(0)"})
    np.savez_compressed(os.path.join(HERE, "case_bench.npz"), **out)
    print({k: getattr(v, "shape", None) for k, v in out.items()})
    for k in ("out_color", "out_depth", "out_transmittance", "out_weight"):
        v = out[k]
        print("   ", k, float(v.min()), float(v.mean()), float(v.max()))


if __name__ == "__main__":
    main()
