"""TEST INFRASTRUCTURE ONLY -- CPU oracle of the training / evaluation image transform that feeds the
MuDPT hot path.  Never imported by the product (mudpt_b200/).

What it restates.  The reference's yaml asks Dassl for
    INPUT.TRANSFORMS = ["random_resized_crop", "random_flip", "normalize"], INTERPOLATION = "bicubic",
    PIXEL_MEAN / PIXEL_STD = CLIP's             (configs/trainers/MuDPT/vit_b16_bz4_ep10_nctx2_depth9.yaml:8-13)
and `MuDPT.parse_batch_train` (trainers/mudpt.py:263-268) receives the resulting float tensor.  Dassl is an
un-vendored, unpinned dependency of the reference; its transform builder maps those names onto
torchvision: RandomResizedCrop(size, scale=(0.08, 1), interpolation=BICUBIC) -> RandomHorizontalFlip() ->
ToTensor() -> Normalize(mean, std) for training and Resize(max(size)) -> CenterCrop(size) -> ToTensor() ->
Normalize for evaluation, applied to PIL images.  The arithmetic therefore lives in
    torchvision 0.26.0 (transforms/transforms.py: RandomResizedCrop.get_params, RandomHorizontalFlip.forward,
                        functional.py: _compute_resized_output_size, center_crop, to_tensor, normalize) and
    Pillow 12.2.0     (libImaging/Resample.c: precompute_coeffs, normalize_coeffs_8bpc,
                        ImagingResampleHorizontal_8bpc / Vertical_8bpc, bicubic_filter with a = -0.5),
both present in this image; the functions below restate their published algorithms.

PINNED: tests/test_input_pipeline.py checks `resample_u8` bit-for-bit against PIL.Image.resize and the whole
pipeline bit-for-bit (fp32) against the torchvision Compose, on random images / boxes, and the RNG-consuming
`random_resized_crop_params` / `random_flip` against torchvision under the same torch seed.
"""
from __future__ import annotations

import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2  # Resample.c: coefficients are fixed point with 22 fractional bits
BICUBIC_SUPPORT = 2.0


def _bicubic(x: float) -> float:
    """Resample.c:bicubic_filter (a = -0.5), evaluated in double precision exactly as the C expression."""
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def precompute_coeffs(in_size: int, in0: float, in1: float, out_size: int):
    """Resample.c:precompute_coeffs + normalize_coeffs_8bpc for the bicubic filter.
    Returns (ksize, bounds [out_size, 2] = (xmin, count), kk [out_size, ksize] int32 fixed-point weights)."""
    scale = filterscale = (in1 - in0) / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = BICUBIC_SUPPORT * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = in0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)  # C cast: truncation toward zero
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = [_bicubic((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            v = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return ksize, bounds, kk


def _clip8(acc: np.ndarray) -> np.ndarray:
    return np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)


def resample_u8(img: np.ndarray, out_w: int, out_h: int, box=None) -> np.ndarray:
    """PIL.Image.resize((out_w, out_h), BICUBIC, box=box) on an 8-bit HWC image: horizontal pass (rounded to
    uint8) then vertical pass, each skipped when it would be the identity (Resample.c:ImagingResample)."""
    H, W = img.shape[:2]
    x0, y0, x1, y1 = box if box is not None else (0, 0, W, H)
    need_h = out_w != W or x0 != 0 or x1 != W
    need_v = out_h != H or y0 != 0 or y1 != H
    cur = img
    if need_h:
        _, bnd, kk = precompute_coeffs(W, x0, x1, out_w)
        out = np.empty((cur.shape[0], out_w, cur.shape[2]), np.uint8)
        for xx in range(out_w):
            xmin, n = bnd[xx]
            acc = (cur[:, xmin:xmin + n, :].astype(np.int64) * kk[xx, :n].astype(np.int64)[None, :, None]).sum(1)
            out[:, xx, :] = _clip8(acc + (1 << (PRECISION_BITS - 1)))
        cur = out
    if need_v:
        _, bnd, kk = precompute_coeffs(H, y0, y1, out_h)
        out = np.empty((out_h, cur.shape[1], cur.shape[2]), np.uint8)
        for yy in range(out_h):
            ymin, n = bnd[yy]
            acc = (cur[ymin:ymin + n].astype(np.int64) * kk[yy, :n].astype(np.int64)[:, None, None]).sum(0)
            out[yy] = _clip8(acc + (1 << (PRECISION_BITS - 1)))
        cur = out
    return cur


def to_tensor_normalize(img_u8: np.ndarray, mean, std) -> np.ndarray:
    """torchvision to_tensor (uint8 HWC -> float32 CHW / 255) followed by normalize ((x - mean) / std), fp32."""
    x = img_u8.astype(np.float32).transpose(2, 0, 1) / np.float32(255)
    m = np.asarray(mean, np.float32)[:, None, None]
    s = np.asarray(std, np.float32)[:, None, None]
    return (x - m) / s


def train_transform(img: np.ndarray, top: int, left: int, h: int, w: int, flip: bool, size, mean, std) -> np.ndarray:
    """RandomResizedCrop (given its drawn box) -> horizontal flip -> ToTensor -> Normalize.
    torchvision crops first (PIL.Image.crop), then resizes the crop: resampling never sees pixels outside the box."""
    crop = img[top:top + h, left:left + w]
    out = resample_u8(crop, size[1], size[0])
    if flip:
        out = out[:, ::-1]
    return to_tensor_normalize(out, mean, std)


def resized_output_size(h: int, w: int, size: int):
    """torchvision functional._compute_resized_output_size for an int size (shorter edge)."""
    short, long = (w, h) if w <= h else (h, w)
    new_short, new_long = size, int(size * long / short)
    new_w, new_h = (new_short, new_long) if w <= h else (new_long, new_short)
    return new_h, new_w


def eval_transform(img: np.ndarray, size, mean, std) -> np.ndarray:
    """Resize(max(size)) -> CenterCrop(size) -> ToTensor -> Normalize (images at least `size` after the resize)."""
    H, W = img.shape[:2]
    nh, nw = resized_output_size(H, W, max(size))
    r = resample_u8(img, nw, nh)
    top = int(round((nh - size[0]) / 2.0))
    left = int(round((nw - size[1]) / 2.0))
    return to_tensor_normalize(r[top:top + size[0], left:left + size[1]], mean, std)
