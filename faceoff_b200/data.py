"""Input pipeline on the GPU.

The reference's loader (TemporalAlignment/dataset.py:235-249,417-420) converts every uint8 frame to a normalised fp32
tensor on the CPU (torchvision ``ToTensor`` + ``Normalize(0.5, 0.5)``), and ``utils.process_data`` (utils.py:29-38)
concatenates the perturbed source-face frames and the background frames on the channel axis and moves fp32 tensors to the
device: 36 bytes per pixel over PCIe.  Here the loader hands over the uint8 frames (9 bytes per pixel, pinned memory, async
copy) and the same arithmetic runs in one kernel per tensor on the device, bit-exact with the torchvision transforms.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import ops
from ._lib import FaceoffB200Error


def _check(name: str, t: torch.Tensor):
    if not (t.dtype == torch.uint8 and t.dim() == 4 and t.shape[-1] == 3):
        raise FaceoffB200Error(f"{name}: uint8 frames [T, H, W, 3] expected, got {t.dtype} {tuple(t.shape)}")


def frames_to_device(frames: torch.Tensor, device, stream: Optional[torch.cuda.Stream] = None) -> torch.Tensor:
    """Asynchronous host -> device copy of uint8 frames (pin the host tensor to make it truly asynchronous)."""
    if stream is None:
        return frames.to(device, non_blocking=True)
    with torch.cuda.stream(stream):
        return frames.to(device, non_blocking=True)


def process_data_u8(source: torch.Tensor, background: torch.Tensor, source_images: torch.Tensor, device=None,
                    mean: float = 0.5, std: float = 0.5, out_img: Optional[torch.Tensor] = None,
                    out_gt: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, int, torch.Tensor]:
    """``utils.process_data`` (utils.py:29-38) for uint8 HWC frames: returns (img [T, 6, H, W], S, ground_truth [T, 3, H, W]),
    fp32 on the device, with img = cat([normalise(source), normalise(background)], 1) and ground_truth =
    normalise(source_images), normalise(x) = ((x / 255) - mean) / std (dataset.py:235-249)."""
    for name, t in (("source", source), ("background", background), ("source_images", source_images)):
        _check(name, t)
    if device is not None:
        source, background, source_images = (t.to(device, non_blocking=True) for t in (source, background, source_images))
    if not source.is_cuda:
        raise FaceoffB200Error("process_data_u8: CUDA tensors (or a device) required; faceoff_b200 has no CPU path")
    t, h, w, _ = source.shape
    img = out_img if out_img is not None else torch.empty((t, 6, h, w), dtype=torch.float32, device=source.device)
    gt = out_gt if out_gt is not None else torch.empty((t, 3, h, w), dtype=torch.float32, device=source.device)
    ops.u8hwc_to_nchw(source.contiguous(), img, 0, mean, std)
    ops.u8hwc_to_nchw(background.contiguous(), img, 3, mean, std)
    ops.u8hwc_to_nchw(source_images.contiguous(), gt, 0, mean, std)
    return img, t, gt
