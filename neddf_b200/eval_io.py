"""Evaluation I/O off the critical path.

``BaseTrainer.render_test`` (neddf/trainer/base_trainer.py:123-174) renders a frame and then, on the same
thread, converts it to uint8 on the host, writes three PNGs and computes PSNR / SSIM - a few hundred
milliseconds per 800x800 frame during which the GPU idles (a frame renders in 2.4 s here, 0.3 s on eight
GPUs).  ``FrameWriter`` keeps the reference's arithmetic and file names but moves it out of the way:

* uint8 conversion on the device (same clamp / scale / truncation as :147-160), on a side stream,
* device->host copy into pinned buffers on that stream,
* PNG encoding, PSNR and (when scikit-image is installed) SSIM in a worker thread,

so the next frame's ``render_image`` starts immediately.  ``render_all`` is the drop-in for
``BaseTrainer.render_all`` (:176-187); ``install(patch_trainer=True)`` binds it.
"""
import math
import queue
import threading
from pathlib import Path
from typing import Dict, Optional

import numpy as np
import torch


def to_uint8_images(images: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """The reference's conversions (base_trainer.py:147-160) as device ops: float -> uint8 truncates like
    ``ndarray.astype(np.uint8)`` does for values already clamped to [0, 255]."""
    rgb = torch.clamp(images["color"] * 255, 0, 255).to(torch.uint8)
    depth = torch.clamp((images["depth"] - 2.0) / 4.0 * 50000 / 256, 0, 255).to(torch.uint8)
    return {"rgb": rgb, "depth": depth}


def psnr_uint8(a: np.ndarray, b: np.ndarray) -> float:
    """skimage.metrics.peak_signal_noise_ratio for uint8 inputs: data_range 255, float64 mean squared error."""
    err = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
    return float("inf") if err == 0 else 10.0 * math.log10((255.0 ** 2) / err)


class FrameWriter:
    def __init__(self, max_pending: int = 2, quiet: bool = False) -> None:
        self._q: "queue.Queue" = queue.Queue(maxsize=max_pending)
        self._quiet = quiet
        self.results = []  # (camera_id, psnr, ssim or None)
        self._err: Optional[BaseException] = None
        self._stream = None
        self._t = threading.Thread(target=self._work, daemon=True)
        self._t.start()

    def submit(self, images: Dict[str, torch.Tensor], rgb_gt: np.ndarray, output_dir: Path, camera_id: int,
               downsampling: int = 1) -> None:
        """Queue one rendered frame (``images`` = render_image's dict, tensors may live on the GPU)."""
        if self._err is not None:
            raise self._err
        dev = images["color"].device
        event = None
        if dev.type == "cuda":
            if self._stream is None:
                self._stream = torch.cuda.Stream(device=dev)
            self._stream.wait_stream(torch.cuda.current_stream(dev))  # the render has been enqueued before us
            with torch.cuda.stream(self._stream):
                u8 = to_uint8_images(images)
                host = {k: torch.empty(v.shape, dtype=torch.uint8, pin_memory=True) for k, v in u8.items()}
                for k in u8:
                    host[k].copy_(u8[k], non_blocking=True)
                    u8[k].record_stream(self._stream)
                for v in images.values():
                    v.record_stream(self._stream)
                event = torch.cuda.Event()
                event.record(self._stream)
        else:
            host = to_uint8_images(images)
        self._q.put((event, host, rgb_gt, Path(output_dir), int(camera_id), int(downsampling)))

    def _work(self) -> None:
        import cv2
        try:
            from skimage.metrics import structural_similarity
        except Exception:  # scikit-image is optional here; the PSNR needs nothing
            structural_similarity = None
        while True:
            item = self._q.get()
            if item is None:
                return
            try:
                event, host, rgb_gt, out, cid, ds = item
                if event is not None:
                    event.synchronize()
                rgb_np, depth_np = host["rgb"].numpy(), host["depth"].numpy()
                cv2.imwrite(str(out / "{:03}_rgb.png".format(cid)), rgb_np)       # base_trainer.py:163-168
                cv2.imwrite(str(out / "{:03}_rgb_gt.png".format(cid)), rgb_gt)
                cv2.imwrite(str(out / "{:03}_depth.png".format(cid)), depth_np)
                if ds == 1:                                                         # :171-174
                    psnr = psnr_uint8(rgb_np, rgb_gt)
                    ssim = structural_similarity(rgb_np, rgb_gt, channel_axis=2) if structural_similarity else None
                    self.results.append((cid, psnr, ssim))
                    if not self._quiet:
                        print("psnr: {}, ssim: {}".format(psnr, ssim))
            except BaseException as e:  # surfaced by the next submit / close
                self._err = e
            finally:
                self._q.task_done()

    def close(self) -> None:
        self._q.put(None)
        self._t.join()
        if self._err is not None:
            raise self._err


def render_all(self, output_dir: Path) -> None:
    """Drop-in for BaseTrainer.render_all (base_trainer.py:176-187): same frames, same files, same prints;
    the conversion / PNG / metric work of frame i overlaps the rendering of frame i + 1."""
    writer = FrameWriter()
    self.neural_render.set_iter(-1)
    try:
        for camera_id in range(len(self.dataset)):
            print("rendering from camera {}".format(camera_id))
            rgb_gt = self.dataset[camera_id]["rgb_images"].astype(np.uint8)
            camera = self.cameras[camera_id]
            camera.update_transform()
            h, w = rgb_gt.shape[0], rgb_gt.shape[1]
            images = self.neural_render.render_image(w, h, camera, ["color", "depth"], 1, self.chunk)
            writer.submit(images, rgb_gt, output_dir, camera_id, 1)
    finally:
        writer.close()
