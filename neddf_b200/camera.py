"""Minimal pinhole camera with the attributes the renderer reads.

The reference's Camera / PinholeCalib (neddf/camera/camera.py, pinhole_calib.py) work with
this renderer unchanged; these stand-ins exist so the package, its tests and bench.py run
without the reference installed.  Only what render_rays / render_image touch is mirrored:
``R`` [3,3], ``T`` [3], ``camera_calib.params`` = [fx, fy, cx, cy], ``device``,
``update_transform()``.
"""
import numpy as np
import torch
from torch import nn


class PinholeCalib(nn.Module):
    def __init__(self, calib_param) -> None:
        super().__init__()
        calib_param = np.asarray(calib_param, dtype=np.float64)
        assert calib_param.shape == (4,)  # pinhole_calib.py: fx, fy, cx, cy
        self.params = nn.Parameter(torch.from_numpy(calib_param).to(torch.float32))

    @property
    def device(self) -> torch.device:
        return self.params.device


def _rodrigues(rotvec: np.ndarray) -> np.ndarray:
    theta = float(np.linalg.norm(rotvec))
    K = np.array([[0, -rotvec[2], rotvec[1]], [rotvec[2], 0, -rotvec[0]], [-rotvec[1], rotvec[0], 0]], dtype=np.float64)
    if theta < 1e-12:
        return np.eye(3) + K
    K = K / theta
    return np.eye(3) + np.sin(theta) * K + (1 - np.cos(theta)) * (K @ K)


class Camera(nn.Module):
    """Fixed-pose camera: R = Rodrigues(param[:3]), T = param[3:6] (camera.py:66-118 with the
    trainable SE(3) delta at zero, which is how the shipped trainer uses it)."""

    def __init__(self, camera_calib: PinholeCalib, initial_camera_param=None) -> None:
        super().__init__()
        if initial_camera_param is None:
            initial_camera_param = np.zeros(6, dtype=np.float32)
        self.camera_calib = camera_calib
        self.initial_params_np = np.asarray(initial_camera_param, dtype=np.float32)
        self.params = nn.Parameter(torch.zeros(6, dtype=torch.float32))
        self.R = torch.eye(3, dtype=torch.float32)
        self.T = torch.zeros(3, dtype=torch.float32)
        self.update_transform()

    @staticmethod
    def from_matrix(camera_calib: PinholeCalib, R, T) -> "Camera":
        cam = Camera(camera_calib)
        cam._fixed = (torch.as_tensor(np.asarray(R), dtype=torch.float32), torch.as_tensor(np.asarray(T), dtype=torch.float32))
        cam.update_transform()
        return cam

    @property
    def device(self) -> torch.device:
        return self.params.device

    def update_transform(self) -> None:
        fixed = getattr(self, "_fixed", None)
        if fixed is not None:
            self.R, self.T = fixed[0].to(self.device), fixed[1].to(self.device)
            return
        R0 = _rodrigues(self.initial_params_np[:3].astype(np.float64))
        self.R = torch.from_numpy(R0.astype(np.float32)).to(self.device)
        self.T = torch.from_numpy(self.initial_params_np[3:6].astype(np.float32)).to(self.device)
