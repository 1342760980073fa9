"""The BENCHMARKED workload against the REAL reference: 1024 random pixels of
bench.py's 800 x 800 frame - its network / render configuration, its seeded weights, its camera 0 - rendered by the
reference's own NeRFRender.render_rays with recorded uniforms (tests/golden/make_bench_golden.py -> case_bench.npz).

CPU: the oracle restatement on that chunk (closes the chain "bench.py's in-run parity block: CUDA path vs oracle" +
"oracle vs real reference" on the benchmarked configuration itself).  GPU: the CUDA path on all three engines."""
import os

import numpy as np
import pytest
import torch

import bench
from oracle import neddf_oracle as orc
from tests.helpers import GOLDEN, PARITY_TOL, nerr


def _load():
    z = np.load(os.path.join(GOLDEN, "case_bench.npz"), allow_pickle=False)
    return {k: z[k] for k in z.files}


def test_oracle_matches_the_reference_on_the_bench_workload():
    z = _load()
    fc, rc = orc.FieldConfig.from_dict(bench.NET_CFG), orc.RenderConfig.from_dict(bench.RENDER_CFG)
    P = bench.seeded_params()
    cam = orc.CameraPose(torch.from_numpy(z["cam_R"]), torch.from_numpy(z["cam_T"]), *[float(v) for v in z["cam_calib"]])
    R, T, calib = bench.synthetic_pose(0)  # the fixture's camera is the bench's camera 0
    assert np.abs(z["cam_R"] - R).max() < 1e-6 and np.abs(z["cam_T"] - T).max() < 1e-6 and np.abs(z["cam_calib"] - calib).max() < 1e-3
    taps = {}
    sl = slice(0, 256)  # a quarter of the chunk keeps the CPU suite short; the GPU test takes all 1024 rays
    with torch.no_grad():
        out = orc.render_rays(P, P, fc, orc.FieldState.at_iter(fc, -1), rc, torch.from_numpy(z["uv"][sl]), cam,
                              torch.from_numpy(z["u_coarse"][sl]), torch.from_numpy(z["u_fine"][sl]), taps=taps)
    assert sorted(out) == sorted(k[4:] for k in z if k.startswith("out_"))
    for k, v in out.items():
        ref = z["out_" + k]
        assert np.abs(v.numpy() - ref[sl]).max() / np.abs(ref).max() < 2e-5, k  # measured 2e-7 .. 5e-6 on the whole chunk
    assert nerr(taps["dists_fine"].numpy(), z["dists_fine"][sl]) < 2e-5


@pytest.mark.gpu
@pytest.mark.parametrize("engine", ["tc", "tc2", "fp32"])
def test_cuda_path_matches_the_reference_on_the_bench_workload(engine):
    import neddf_b200
    dev = torch.device("cuda:0")
    z = _load()
    sd, _ = bench.seeded_state_dict()
    render = neddf_b200.NeRFRender(network_config=bench.NET_CFG, **bench.RENDER_CFG)
    render.load_state_dict(sd)
    render.to(dev)
    render.set_iter(-1)
    render.set_engine(engine)
    cam = neddf_b200.Camera.from_matrix(neddf_b200.PinholeCalib(z["cam_calib"]), z["cam_R"], z["cam_T"]).to(dev)
    cam.update_transform()
    with torch.no_grad():
        out = render.render_rays(torch.from_numpy(z["uv"]).to(dev), cam,
                                 uniforms=(torch.from_numpy(z["u_coarse"]).to(dev), torch.from_numpy(z["u_fine"]).to(dev)))
    assert sorted(out) == sorted(k[4:] for k in z if k.startswith("out_"))
    for k, v in out.items():
        tol = 1e-3 if k == "weight" else PARITY_TOL  # same bounds as test_render_rays_matches_reference (tests/test_arbiter.py)
        assert nerr(v.cpu().numpy(), z["out_" + k]) < tol, (engine, k)
