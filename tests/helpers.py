"""Shared test helpers: golden-case loading and the parity metric."""
import json
import os

import numpy as np
import torch

from oracle import neddf_oracle as orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# tensor-normalised parity bound of BASELINE.json north_star ("<=1e-4 rel fp32")
PARITY_TOL = 1e-4


def nerr(new, ref) -> float:
    """max|new-ref| / max|ref| (the SURVEY 8(d) parity metric)."""
    new = np.asarray(new, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert new.shape == ref.shape, (new.shape, ref.shape)
    den = max(float(np.abs(ref).max()), 1e-30)
    return float(np.abs(new - ref).max() / den)


def assert_parity(new, ref, tol, kinked=False, what=""):
    """Parity assertion.  ``kinked`` = the configuration uses a piecewise-linear hidden
    activation (ReLU / LeakyReLU): its Jacobian rows are discontinuous where a pre-activation
    crosses 0, so two fp32 evaluations that differ only in summation order legitimately
    disagree by O(1e-4..1e-3) on the handful of samples that sit within rounding distance of
    a kink (the reference disagrees with its own fp64 run there).  For those configs the bound
    is applied to all but 1% of the elements (at least 2) and the outliers are capped at 5e-2."""
    new = np.asarray(new, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert new.shape == ref.shape, (what, new.shape, ref.shape)
    den = max(float(np.abs(ref).max()), 1e-30)
    err = np.abs(new - ref) / den
    if not kinked:
        assert float(err.max()) < tol, (what, float(err.max()))
        return
    n_out = int((err >= tol).sum())
    assert n_out <= max(2, int(1e-2 * err.size)) and float(err.max()) < 5e-2, (what, n_out, err.size, float(err.max()))


class Case:
    """One golden case: configs, weights, camera, inputs and the reference's outputs."""

    def __init__(self, name: str):
        z = np.load(os.path.join(GOLDEN, f"case_{name}.npz"), allow_pickle=False)
        self.z = {k: z[k] for k in z.files}
        meta = json.loads(str(self.z["cfg"]))
        self.net_cfg = meta["network"]
        self.render_cfg = meta["render"]
        self.fc = orc.FieldConfig.from_dict(self.net_cfg)
        self.rc = orc.RenderConfig.from_dict(self.render_cfg)
        w = meta["weights"]
        if isinstance(w, str):
            wz = np.load(os.path.join(GOLDEN, w))
            pf = {k: torch.from_numpy(wz[k]) for k in wz.files}
            pc = pf
        else:
            pf = orc.init_params(self.fc, w["seed"], w["bias_std"])
            pc = orc.init_params(self.fc, w["seed"] + 1, w["bias_std"]) if w["separate"] else pf
        self.p_fine, self.p_coarse = pf, pc
        self.separate = pc is not pf
        calib = self.z["cam_calib"]
        self.cam = orc.CameraPose(torch.from_numpy(self.z["cam_R"]), torch.from_numpy(self.z["cam_T"]),
                                  float(calib[0]), float(calib[1]), float(calib[2]), float(calib[3]))
        self.iter = int(self.z["iter"]) if "iter" in self.z else -1
        self.st = orc.FieldState.at_iter(self.fc, self.iter)
        self.kinked = self.fc.activation_type in ("ReLU", "LeakyReLU")

    def state_dict(self):
        sd = {"network_fine." + k: v for k, v in self.p_fine.items()}
        sd.update({"network_coarse." + k: v for k, v in self.p_coarse.items()})
        return sd

    def t(self, key):
        return torch.from_numpy(self.z[key])

    def outputs(self):
        return {k[4:]: v for k, v in self.z.items() if k.startswith("out_")}
