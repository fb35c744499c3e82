"""GPU input pipeline in front of the vision tower.

Host side of `mudpt_augment_images` (include/mudpt_b200.h, csrc/augment.cu).  It stands where the reference
has Dassl's transform builder applied by the data-loader workers to PIL images (un-vendored dependency; the
yaml asks for INPUT.TRANSFORMS = random_resized_crop, random_flip, normalize with INTERPOLATION = bicubic and
CLIP's PIXEL_MEAN / PIXEL_STD, configs/trainers/MuDPT/vit_b16_bz4_ep10_nctx2_depth9.yaml:8-13), i.e.
torchvision's RandomResizedCrop -> RandomHorizontalFlip -> ToTensor -> Normalize for training and
Resize(max(size)) -> CenterCrop(size) -> ToTensor -> Normalize for evaluation.  The batch it returns is the
`batch["img"]` tensor of `MuDPT.parse_batch_train` (trainers/mudpt.py:263-268).

Only the random draws happen here (same torch CPU generator calls, in the same order, as torchvision's
`RandomResizedCrop.get_params` and `RandomHorizontalFlip.forward`, so a seeded run crops the same boxes as the
reference's loader would); every pixel is produced by the CUDA kernels, bit-identical to the PIL / torchvision
result.  JPEG decoding is not part of this module: it takes decoded 8-bit RGB images (HWC).  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib

CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)

# mudpt_image_desc (include/mudpt_b200.h), 64 bytes
DESC_DTYPE = np.dtype([("src", "<u8"), ("height", "<i4"), ("width", "<i4"), ("pitch", "<i4"),
                       ("box_x", "<i4"), ("box_y", "<i4"), ("box_w", "<i4"), ("box_h", "<i4"),
                       ("rs_w", "<i4"), ("rs_h", "<i4"), ("win_x", "<i4"), ("win_y", "<i4"),
                       ("flip", "<i4"), ("reserved", "<i4", (2,))])
assert DESC_DTYPE.itemsize == 64


def random_resized_crop_params(height: int, width: int, scale=(0.08, 1.0), ratio=(3.0 / 4.0, 4.0 / 3.0)):
    """(top, left, h, w) drawn as torchvision 0.26 RandomResizedCrop.get_params does (same generator calls)."""
    area = height * width
    log_ratio = torch.log(torch.tensor(ratio))
    for _ in range(10):
        target_area = area * torch.empty(1).uniform_(scale[0], scale[1]).item()
        aspect_ratio = torch.exp(torch.empty(1).uniform_(log_ratio[0], log_ratio[1])).item()
        w = int(round(math.sqrt(target_area * aspect_ratio)))
        h = int(round(math.sqrt(target_area / aspect_ratio)))
        if 0 < w <= width and 0 < h <= height:
            i = torch.randint(0, height - h + 1, size=(1,)).item()
            j = torch.randint(0, width - w + 1, size=(1,)).item()
            return i, j, h, w
    in_ratio = float(width) / float(height)  # fall back to a central crop
    if in_ratio < min(ratio):
        w = width
        h = int(round(w / min(ratio)))
    elif in_ratio > max(ratio):
        h = height
        w = int(round(h * max(ratio)))
    else:
        w, h = width, height
    return (height - h) // 2, (width - w) // 2, h, w


def random_flip(p: float = 0.5) -> bool:
    """torchvision RandomHorizontalFlip.forward's draw."""
    return bool(torch.rand(1) < p)


def resized_output_size(height: int, width: int, size: int) -> Tuple[int, int]:
    """torchvision Resize(int): the shorter edge becomes `size` (functional._compute_resized_output_size)."""
    short, long = (width, height) if width <= height else (height, width)
    new_short, new_long = size, int(size * long / short)
    new_w, new_h = (new_short, new_long) if width <= height else (new_long, new_short)
    return new_h, new_w


def draw_geometry(height: int, width: int, size, is_train: bool, scale=(0.08, 1.0), ratio=(3.0 / 4.0, 4.0 / 3.0),
                  flip_p: float = 0.5):
    """Per-image geometry (box_x, box_y, box_w, box_h, rs_w, rs_h, win_x, win_y, flip) of mudpt_image_desc.
    Training: the random crop resampled to `size`, then the flip draw (Compose order).  Evaluation: the whole
    image resampled so that its shorter edge is max(size), centre window (torchvision center_crop rounding)."""
    oh, ow = size
    if is_train:
        top, left, h, w = random_resized_crop_params(height, width, scale, ratio)
        flip = random_flip(flip_p) if flip_p > 0 else False
        return left, top, w, h, ow, oh, 0, 0, int(flip)
    nh, nw = resized_output_size(height, width, max(size))
    if nh < oh or nw < ow:
        raise ValueError("evaluation transform: image smaller than the crop after Resize (padding not supported)")
    return 0, 0, width, height, nw, nh, int(round((nw - ow) / 2.0)), int(round((nh - oh) / 2.0)), 0


class GpuTransform:
    """Drop-in for the transform the reference's yaml names, on the GPU.

    transform(images) -> float32 [B, 3, size[0], size[1]] on `device`; `images` are uint8 [H, W, 3] tensors of
    any sizes (CUDA, or CPU -- pinned for an asynchronous upload)."""

    SUPPORTED_TRAIN = ("random_resized_crop", "random_flip", "normalize")

    def __init__(self, size=(224, 224), is_train: bool = True, scale=(0.08, 1.0), ratio=(3.0 / 4.0, 4.0 / 3.0),
                 flip_p: float = 0.5, mean=CLIP_MEAN, std=CLIP_STD, interpolation: str = "bicubic", device=None):
        if interpolation != "bicubic":
            raise NotImplementedError(f"interpolation {interpolation!r}: the reference's configs use bicubic only")
        self.size = (int(size[0]), int(size[1]))
        self.is_train, self.scale, self.ratio, self.flip_p = is_train, tuple(scale), tuple(ratio), flip_p
        self.mean = (C.c_float * 3)(*mean)
        self.std = (C.c_float * 3)(*std)
        self.device = torch.device(device if device is not None else "cuda")
        if self.device.type != "cuda":
            raise RuntimeError("mudpt_b200.input_pipeline runs on a CUDA device only (no CPU fallback)")
        self._lib = _lib.load()
        self._workspace: Optional[torch.Tensor] = None
        self._desc_host: Optional[torch.Tensor] = None

    @classmethod
    def from_cfg(cls, cfg, is_train: bool, device=None) -> "GpuTransform":
        """cfg.INPUT.{SIZE, INTERPOLATION, PIXEL_MEAN, PIXEL_STD, TRANSFORMS} as in the reference's yaml."""
        inp = cfg.INPUT
        names = tuple(getattr(inp, "TRANSFORMS", cls.SUPPORTED_TRAIN))
        for n in names:
            if n not in cls.SUPPORTED_TRAIN:
                raise NotImplementedError(f"transform {n!r} is not used by the reference's MuDPT configs")
        scale = tuple(getattr(inp, "RRCROP_SCALE", (0.08, 1.0)))
        return cls(size=tuple(inp.SIZE), is_train=is_train, scale=scale, mean=tuple(getattr(inp, "PIXEL_MEAN", CLIP_MEAN)),
                   std=tuple(getattr(inp, "PIXEL_STD", CLIP_STD)), interpolation=getattr(inp, "INTERPOLATION", "bicubic"),
                   device=device, flip_p=0.5 if "random_flip" in names else 0.0)

    def draw(self, height: int, width: int):
        return draw_geometry(height, width, self.size, self.is_train, self.scale, self.ratio, self.flip_p)

    def describe(self, images: Sequence[torch.Tensor], params=None) -> np.ndarray:
        """mudpt_image_desc array for device-resident images; `params` overrides the random draws."""
        d = np.zeros(len(images), DESC_DTYPE)
        for i, im in enumerate(images):
            if im.dtype != torch.uint8 or im.dim() != 3 or im.shape[2] != 3 or not im.is_cuda or im.stride(2) != 1 or \
                    im.stride(1) != 3:
                raise ValueError("images must be CUDA uint8 [H, W, 3] tensors with packed pixels")
            H, W = int(im.shape[0]), int(im.shape[1])
            g = params[i] if params is not None else self.draw(H, W)
            d[i] = (im.data_ptr(), H, W, int(im.stride(0)), *[int(v) for v in g], (0, 0))
        return d

    # ------------------------------------------------------------------ the call
    def __call__(self, images: Sequence[torch.Tensor], params=None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        dev = self.device
        imgs: List[torch.Tensor] = [im if im.is_cuda else im.to(dev, non_blocking=True) for im in images]
        n = len(imgs)
        if n == 0:
            raise ValueError("empty batch")
        desc = self.describe(imgs, params)
        oh, ow = self.size
        host_ptr = desc.ctypes.data_as(C.c_void_p)
        need = self._lib.mudpt_augment_workspace_bytes(host_ptr, n, oh, ow)
        if need < 0:
            _lib.check(-1)
        if self._workspace is None or self._workspace.numel() < need:
            self._workspace = torch.empty(int(need), dtype=torch.uint8, device=dev)
        if self._desc_host is None or self._desc_host.numel() < desc.nbytes:
            self._desc_host = torch.empty(max(desc.nbytes, 64 * 64), dtype=torch.uint8).pin_memory()
            self._desc_event = torch.cuda.Event()
        else:
            self._desc_event.synchronize()  # the previous batch's descriptor upload has left the pinned buffer
        self._desc_host[:desc.nbytes].copy_(torch.from_numpy(desc.view(np.uint8).reshape(-1)))
        desc_dev = self._desc_host[:desc.nbytes].to(dev, non_blocking=True)
        self._desc_event.record(torch.cuda.current_stream(dev))
        if out is None:
            out = torch.empty(n, 3, oh, ow, dtype=torch.float32, device=dev)
        assert out.is_cuda and out.is_contiguous() and out.dtype == torch.float32 and tuple(out.shape) == (n, 3, oh, ow)
        _lib.check(self._lib.mudpt_augment_images(desc_dev.data_ptr(), host_ptr, n, oh, ow, self.mean, self.std,
                                                  self._workspace.data_ptr(), self._workspace.numel(), out.data_ptr(),
                                                  _lib.stream_ptr(dev)))
        self._keepalive = (imgs, desc_dev)  # until the next call: the kernels read them asynchronously
        return out
