import numpy as np


def peak_signal_noise_ratio(a, b, data_range=255):
    mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
    return float(10 * np.log10(data_range ** 2 / mse)) if mse > 0 else float("inf")


def structural_similarity(a, b, channel_axis=None, **kw):
    raise NotImplementedError("harness stand-in: SSIM needs scikit-image")
