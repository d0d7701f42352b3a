"""CPU oracle of the ViLT image pre-processing path.  TEST INFRASTRUCTURE.

``reference_preprocess`` follows transformers==4.48.0 ``ViltImageProcessor.preprocess`` (HF:models/vilt/image_processing_vilt.py) step by
step with the same third-party calls the reference makes -- Pillow's ``Image.resize(BICUBIC, reducing_gap=None)`` on uint8, numpy float64
rescale cast to float32, float32 normalise -- followed by the bottom/right zero padding and pixel mask of HF ``pad`` / the reference's
``safe_dict_concat`` (ref:vault/vl_utils/dataset_utils.py:7-36).  ``pillow_resample_u8`` is an independent pure-numpy restatement of Pillow's
ImagingResample (src/libImaging/Resample.c: precompute_coeffs, normalize_coeffs_8bpc, horizontal then vertical 8bpc passes) that
tests/test_image_processing.py pins against Pillow itself, bit for bit.  (The installed transformers 5.5 processor is torchvision-based and
differs from the pinned 4.48.0 / Pillow path by +-1 LSB: it is not the target.)"""
from __future__ import annotations

import math
from typing import List, Sequence, Tuple

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def output_size(h: int, w: int, shorter: int = 384, longer: int = None, size_divisor: int = 32) -> Tuple[int, int]:
    """get_resize_output_image_size of 4.48.0 (min_size/max_size formulation).  NB: longer = int(1333 / 800 * 384) = 639 (639.84 truncated),
    so after the floor to a multiple of 32 the longest side this processor ever emits is 608, not 640."""
    if longer is None:
        longer = int(1333 / 800 * shorter)
    min_size, max_size = shorter, longer
    scale = min_size / min(h, w)
    if h < w:
        newh, neww = min_size, scale * w
    else:
        newh, neww = scale * h, min_size
    if max(newh, neww) > max_size:
        scale = max_size / max(newh, neww)
        newh = newh * scale
        neww = neww * scale
    newh, neww = int(newh + 0.5), int(neww + 0.5)
    return newh // size_divisor * size_divisor, neww // size_divisor * size_divisor


def _bicubic_filter(x: float) -> float:
    a = -0.5
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def _coeffs(in_size: int, out_size: int):
    filterscale = scale = in_size / out_size
    filterscale = max(filterscale, 1.0)
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    out = []
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        k = [_bicubic_filter((x + xmin - center + 0.5) * (1.0 / filterscale)) for x in range(xmax)]
        ww = 0.0
        for v in k:
            ww += v
        if ww != 0.0:
            k = [v / ww for v in k]
        kk = [int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS)) for v in k]
        out.append((xmin, np.asarray(kk, dtype=np.int64)))
    return out


def _pass(img: np.ndarray, out_size: int, axis: int) -> np.ndarray:
    """One 8-bit resampling pass along `axis` (0 = vertical, 1 = horizontal) of an HWC uint8 image."""
    src = np.moveaxis(img, axis, 0).astype(np.int64)
    res = np.empty((out_size,) + src.shape[1:], dtype=np.uint8)
    for xx, (xmin, kk) in enumerate(_coeffs(src.shape[0], out_size)):
        acc = (1 << (PRECISION_BITS - 1)) + np.tensordot(kk, src[xmin:xmin + len(kk)], axes=(0, 0))
        res[xx] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(res, 0, axis)


def pillow_resample_u8(img: np.ndarray, out_hw: Tuple[int, int]) -> np.ndarray:
    """Pillow ImagingResample for an 8-bit RGB image: horizontal pass (if the width changes), then vertical (if the height changes)."""
    h, w = out_hw
    if img.shape[1] != w:
        img = _pass(img, w, 1)
    if img.shape[0] != h:
        img = _pass(img, h, 0)
    return img


def reference_preprocess(images: Sequence[np.ndarray], shorter: int = 384, size_divisor: int = 32, use_pillow: bool = True):
    """-> (pixel_values float32 [B,3,Hmax,Wmax], pixel_mask int64 [B,Hmax,Wmax], list of resized uint8 HWC images)."""
    from PIL import Image

    longer = int(1333 / 800 * shorter)
    resized: List[np.ndarray] = []
    for im in images:
        ho, wo = output_size(im.shape[0], im.shape[1], shorter, longer, size_divisor)
        if use_pillow:
            r = np.array(Image.fromarray(im).resize((wo, ho), resample=Image.BICUBIC, reducing_gap=None))
        else:
            r = pillow_resample_u8(im, (ho, wo))
        resized.append(r)
    Hm, Wm = max(r.shape[0] for r in resized), max(r.shape[1] for r in resized)
    pv = np.zeros((len(images), 3, Hm, Wm), dtype=np.float32)
    pm = np.zeros((len(images), Hm, Wm), dtype=np.int64)
    for n, r in enumerate(resized):
        x = (r.astype(np.float64) * (1 / 255)).astype(np.float32)                        # rescale(image, 1/255, dtype=float32)
        x = (x - np.array([0.5, 0.5, 0.5], dtype=np.float32)) / np.array([0.5, 0.5, 0.5], dtype=np.float32)  # normalize, channels last
        pv[n, :, :r.shape[0], :r.shape[1]] = x.transpose(2, 0, 1)
        pm[n, :r.shape[0], :r.shape[1]] = 1
    return pv, pm, resized
