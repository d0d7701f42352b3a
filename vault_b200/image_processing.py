"""ViLT image pre-processing on the GPU.

Drop-in for what ``ViltImageProcessor`` (transformers==4.48.0, HF:models/vilt/image_processing_vilt.py:37-38, 55-62, 87-113) does to the
images of a batch, and for the zero-padding / mask of ``safe_dict_concat`` (ref:vault/vl_utils/dataset_utils.py:7-36) that follows it in
the reference's loaders:

    resize (shorter side -> 384, longer side <= int(1333/800*384) = 640, both floored to multiples of 32; PIL BICUBIC on uint8)
    -> rescale 1/255 -> normalise (mean = std = 0.5) -> zero-pad to the batch maximum -> pixel_mask

The resize is Pillow's fixed-point two-pass resampler; this module computes its filter taps on the host exactly as Pillow does
(``precompute_coeffs`` + ``normalize_coeffs_8bpc`` of src/libImaging/Resample.c, double precision -> 22-bit fixed point) and the
kernels of csrc/image_prep.cu apply them in integer arithmetic, so the uint8 image -- and with the 256-entry table the fp32 output --
is bit-identical to the CPU pipeline.  No CPU fallback: the arithmetic on pixels happens on the device only.
"""
from __future__ import annotations

import functools
import math
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch

from . import _abi

PRECISION_BITS = 32 - 8 - 2
MAX_LONGER_EDGE, MAX_SHORTER_EDGE = 1333, 800

DESC_DTYPE = np.dtype([("src_off", "<i8"), ("tmp_off", "<i8"), ("h_in", "<i4"), ("w_in", "<i4"), ("h_out", "<i4"), ("w_out", "<i4"),
                       ("ksize_h", "<i4"), ("ksize_v", "<i4"), ("coef_h", "<i4"), ("bound_h", "<i4"), ("coef_v", "<i4"), ("bound_v", "<i4")])


def resize_output_size(h: int, w: int, shorter: int = 384, size_divisor: int = 32) -> Tuple[int, int]:
    """HF:models/vilt/image_processing_vilt.py get_resize_output_image_size (4.48.0): shorter side -> `shorter`, longer side capped at
    int(1333/800 * shorter), round half up, floor to a multiple of `size_divisor`."""
    longer = int(MAX_LONGER_EDGE / MAX_SHORTER_EDGE * shorter)
    scale = shorter / min(h, w)
    if h < w:
        nh, nw = shorter, scale * w
    else:
        nh, nw = scale * h, shorter
    if max(nh, nw) > longer:
        scale = longer / max(nh, nw)
        nh, nw = nh * scale, nw * scale
    nh, nw = int(nh + 0.5), int(nw + 0.5)
    return nh // size_divisor * size_divisor, nw // size_divisor * size_divisor


def _bicubic(x: float) -> float:
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


@functools.lru_cache(maxsize=512)
def pillow_bicubic_taps(in_size: int, out_size: int) -> Tuple[np.ndarray, np.ndarray, int]:
    """Pillow's precompute_coeffs(inSize, 0, inSize, outSize, BICUBIC) + normalize_coeffs_8bpc: (bounds int32 [out,2] = first source
    index and tap count, coefs int32 [out, ksize] in 22-bit fixed point, ksize)."""
    scale = filterscale = float(in_size) / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    coefs = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
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
        if ww != 0.0:
            w = [v / ww for v in w]
        for x, v in enumerate(w):
            coefs[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx, 0], bounds[xx, 1] = xmin, xmax
    return bounds, coefs, ksize


def _to_hwc_uint8(img) -> np.ndarray:
    if isinstance(img, torch.Tensor):
        img = img.detach().cpu().numpy()
    elif not isinstance(img, np.ndarray):  # PIL.Image
        img = np.asarray(img.convert("RGB"))
    if img.dtype != np.uint8 or img.ndim != 3 or img.shape[2] != 3:
        raise ValueError(f"image must be uint8 HWC RGB, got {img.dtype} {img.shape}")
    return np.ascontiguousarray(img)


class ViltImageProcessorB200:
    """``processor(images)`` -> {"pixel_values": fp32 [B,3,Hmax,Wmax] (cuda), "pixel_mask": int64 [B,Hmax,Wmax] (cuda)}."""

    def __init__(self, shortest_edge: int = 384, size_divisor: int = 32, image_mean: Sequence[float] = (0.5, 0.5, 0.5),
                 image_std: Sequence[float] = (0.5, 0.5, 0.5), rescale_factor: float = 1 / 255, device="cuda"):
        self.shortest_edge, self.size_divisor, self.device = shortest_edge, size_divisor, torch.device(device)
        v = (np.arange(256).astype(np.float64) * rescale_factor).astype(np.float32)  # HF rescale: float64 product, cast to float32
        lut = np.stack([(v - np.float32(m)) / np.float32(s) for m, s in zip(image_mean, image_std)]).astype(np.float32)  # HF normalize in float32
        self._lut_host = lut
        self._lut = None

    def plan(self, sizes: Sequence[Tuple[int, int]]) -> Dict[str, object]:
        """Host side of a batch: output sizes, tap tables, descriptors (pure integer/double bookkeeping, no pixel is touched)."""
        descs = np.zeros(len(sizes), dtype=DESC_DTYPE)
        coef_parts: List[np.ndarray] = []
        bound_parts: List[np.ndarray] = []
        table_off: Dict[Tuple[int, int], Tuple[int, int, int]] = {}
        n_coef = n_bound = 0
        src_off = tmp_off = 0

        def table(i, o):
            nonlocal n_coef, n_bound
            if (i, o) not in table_off:
                b, c, k = pillow_bicubic_taps(i, o)
                table_off[(i, o)] = (n_coef, n_bound, k)
                coef_parts.append(c.reshape(-1))
                bound_parts.append(b.reshape(-1))
                n_coef += c.size
                n_bound += b.size
            return table_off[(i, o)]

        for n, (h, w) in enumerate(sizes):
            ho, wo = resize_output_size(h, w, self.shortest_edge, self.size_divisor)
            ch, bh, kh = table(w, wo)
            cv, bv, kv = table(h, ho)
            descs[n] = (src_off, tmp_off, h, w, ho, wo, kh, kv, ch, bh, cv, bv)
            src_off += h * w * 3
            tmp_off += h * wo * 3
        return dict(descs=descs, coefs=np.concatenate(coef_parts), bounds=np.concatenate(bound_parts), src_bytes=src_off, tmp_bytes=tmp_off,
                    Hmax=int(descs["h_out"].max()), Wmax=int(descs["w_out"].max()), max_h_in=int(descs["h_in"].max()), max_w_out=int(descs["w_out"].max()))

    def __call__(self, images, return_tensors: str = "pt") -> Dict[str, torch.Tensor]:
        if self.device.type != "cuda":
            raise RuntimeError("vault_b200 image pre-processing runs on CUDA only -- there is no CPU path")
        if not isinstance(images, (list, tuple)):
            images = [images]
        arrs = [_to_hwc_uint8(im) for im in images]
        p = self.plan([a.shape[:2] for a in arrs])
        dev = self.device
        src = torch.empty(p["src_bytes"], dtype=torch.uint8).pin_memory()
        flat = src.numpy()
        off = 0
        for a in arrs:
            flat[off:off + a.size] = a.reshape(-1)
            off += a.size
        if self._lut is None or self._lut.device != dev:
            self._lut = torch.from_numpy(self._lut_host).to(dev)
        src_d = src.to(dev, non_blocking=True)
        descs_d = torch.from_numpy(p["descs"].view(np.uint8).copy()).to(dev, non_blocking=True)
        coefs_d = torch.from_numpy(p["coefs"]).to(dev, non_blocking=True)
        bounds_d = torch.from_numpy(p["bounds"]).to(dev, non_blocking=True)
        tmp = torch.empty(max(p["tmp_bytes"], 1), dtype=torch.uint8, device=dev)
        B, Hm, Wm = len(arrs), p["Hmax"], p["Wmax"]
        pixel_values = torch.empty((B, 3, Hm, Wm), dtype=torch.float32, device=dev)
        pixel_mask = torch.empty((B, Hm, Wm), dtype=torch.int64, device=dev)
        _abi.call("vault_image_preprocess", src_d.data_ptr(), descs_d.data_ptr(), coefs_d.data_ptr(), bounds_d.data_ptr(), tmp.data_ptr(),
                  self._lut.data_ptr(), pixel_values.data_ptr(), pixel_mask.data_ptr(), B, Hm, Wm, p["max_h_in"], p["max_w_out"],
                  torch.cuda.current_stream(dev).cuda_stream)
        return {"pixel_values": pixel_values, "pixel_mask": pixel_mask}
