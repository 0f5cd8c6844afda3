"""numpy restatement of the camera-mode preprocessing of the reference  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

What it restates (file:line relative to the reference tree):
  functions/functions_RESNET50_Truncate_Gram_Attention.py:499-501   cv2.cvtColor(frame, BGR2RGB) -> Image.fromarray
                                                                     -> transform(pil_img)
  test_RESNET50_Truncate_gram_attention.py:61-66                    transforms.Compose([Resize, CenterCrop, ToTensor,
                                                                     Normalize(mean, std)])
The arithmetic lives in third-party code that is not under /root/reference: torchvision.transforms (0.26.0 here;
unpinned in requirements.txt:10-11) calls PIL.Image.resize(..., BILINEAR), i.e. Pillow's ImagingResample
(src/libImaging/Resample.c; Pillow 12.2.0 here): a separable triangle filter whose support grows with the down-scale
factor (antialiasing), evaluated in 8-bit fixed point: coefficients rounded to 22 fractional bits, a horizontal pass
that rounds to uint8, then a vertical pass that rounds to uint8. ToTensor divides by 255 in float32 and Normalize
computes (x - mean) / std in float32.

Pinning: tests/test_preprocess_oracle.py checks this restatement for exact equality against Pillow / torchvision
themselves (both are installed wherever the tests run).
"""
from __future__ import annotations

import math
from typing import Sequence, Tuple

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def precompute_coeffs(in_size: int, out_size: int):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc for the bilinear filter over the full axis.
    Returns (xmin[out], xsize[out], kk[out, kmax]) with integer coefficients."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    kmax = int(math.ceil(support)) * 2 + 1
    xmin = np.zeros(out_size, dtype=np.int32)
    xsize = np.zeros(out_size, dtype=np.int32)
    kk = np.zeros((out_size, kmax), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        lo = int(center - support + 0.5)
        lo = max(lo, 0)
        hi = int(center + support + 0.5)
        hi = min(hi, in_size)
        n = hi - lo
        w = np.array([max(0.0, 1.0 - abs((x + lo - center + 0.5) * ss)) for x in range(n)], dtype=np.float64)
        tot = w.sum()
        if tot != 0.0:
            w = w / tot
        xmin[xx], xsize[xx] = lo, n
        for x in range(n):
            kk[xx, x] = int(w[x] * (1 << PRECISION_BITS) + 0.5) if w[x] >= 0 else int(-0.5 + w[x] * (1 << PRECISION_BITS))
    return xmin, xsize, kk


def _clip8(acc: np.ndarray) -> np.ndarray:
    return np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)


def resize_bilinear_u8(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """(H, W, 3) uint8 -> (out_h, out_w, 3) uint8, as PIL.Image.resize((out_w, out_h), BILINEAR) does."""
    h, w, _ = img.shape
    src = img.astype(np.int64)
    if out_w != w:
        xmin, xsize, kk = precompute_coeffs(w, out_w)
        tmp = np.zeros((h, out_w, 3), dtype=np.uint8)
        for xx in range(out_w):
            acc = np.full((h, 3), 1 << (PRECISION_BITS - 1), dtype=np.int64)
            for x in range(xsize[xx]):
                acc += src[:, xmin[xx] + x, :] * int(kk[xx, x])
            tmp[:, xx, :] = _clip8(acc)
        src = tmp.astype(np.int64)
        w = out_w
    if out_h != h:
        ymin, ysize, kk = precompute_coeffs(h, out_h)
        out = np.zeros((out_h, w, 3), dtype=np.uint8)
        for yy in range(out_h):
            acc = np.full((w, 3), 1 << (PRECISION_BITS - 1), dtype=np.int64)
            for y in range(ysize[yy]):
                acc += src[ymin[yy] + y, :, :] * int(kk[yy, y])
            out[yy] = _clip8(acc)
        return out
    return src.astype(np.uint8)


def resize_output_size(h: int, w: int, size) -> Tuple[int, int]:
    """torchvision.transforms.Resize: an int resizes the shorter side to it and keeps the aspect ratio."""
    if isinstance(size, (tuple, list)):
        if len(size) == 2:
            return int(size[0]), int(size[1])
        size = size[0]
    short, long_ = (w, h) if w <= h else (h, w)
    new_short, new_long = int(size), int(size * long_ / short)
    return (new_long, new_short) if w <= h else (new_short, new_long)


def center_crop_box(h: int, w: int, size) -> Tuple[int, int, int, int]:
    """torchvision.transforms.functional.center_crop for an image at least as large as the crop: (top, left, th, tw)."""
    th, tw = (int(size), int(size)) if not isinstance(size, (tuple, list)) else (int(size[0]), int(size[-1]))
    return int(round((h - th) / 2.0)), int(round((w - tw) / 2.0)), th, tw


def camera_preprocess(frame_bgr: np.ndarray, resize, crop, mean: Sequence[float], std: Sequence[float]) -> np.ndarray:
    """BGR uint8 frame (H, W, 3) -> normalised float32 (3, h, w): cvtColor(BGR2RGB), Resize(resize), optional
    CenterCrop(crop), ToTensor, Normalize(mean, std)."""
    rgb = frame_bgr[:, :, ::-1]
    oh, ow = resize_output_size(rgb.shape[0], rgb.shape[1], resize)
    img = resize_bilinear_u8(np.ascontiguousarray(rgb), oh, ow)
    if crop is not None:
        top, left, th, tw = center_crop_box(oh, ow, crop)
        img = img[top:top + th, left:left + tw]
    t = img.astype(np.float32).transpose(2, 0, 1) / np.float32(255.0)
    m = np.asarray(mean, dtype=np.float32).reshape(3, 1, 1)
    s = np.asarray(std, dtype=np.float32).reshape(3, 1, 1)
    return ((t - m) / s).astype(np.float32)
