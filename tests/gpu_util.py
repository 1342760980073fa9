"""Helpers for the GPU parity tests (imported only by tests marked gpu)."""
import torch

import neddf_b200
from tests.helpers import Case

DEV = torch.device("cuda:0")


def build_render(c: Case, engine: str = "fp32") -> neddf_b200.NeRFRender:
    r = neddf_b200.NeRFRender(network_config=c.net_cfg, **{k: v for k, v in c.render_cfg.items() if k != "_target_"})
    missing = r.load_state_dict(c.state_dict())
    assert not missing.missing_keys and not missing.unexpected_keys
    r.to(DEV)
    r.set_iter(c.iter)
    r.set_engine(engine)
    return r


def build_camera(c: Case) -> neddf_b200.Camera:
    cam = neddf_b200.Camera.from_matrix(neddf_b200.PinholeCalib(c.z["cam_calib"]), c.z["cam_R"], c.z["cam_T"]).to(DEV)
    cam.update_transform()
    return cam
