"""Restatement of Pillow's `Image.resize(size)` (default BICUBIC, antialiased) -- TEST INFRASTRUCTURE.

The reference preprocesses every System-1 frame with `Image.fromarray(x).resize((224, 224))`
(internnav/agent/internvla_n1_agent.py L309-321): Pillow's two-pass separable resampler, src/libImaging/Resample.c
(Pillow is a dependency of the reference, not part of its tree; algorithm restated from its published source):

  precompute_coeffs   per output index: centre, window [xmin, xmax) of width <= ceil(support) * 2 + 1 with
                      support = 2 * max(scale, 1), bicubic weights (a = -0.5) normalised to sum 1;
  8-bit images        weights -> fixed point (22 fractional bits, round half away from zero), int32 accumulate starting
                      at 1 << 21, arithmetic shift, clamp to [0, 255]; the horizontal pass output is 8-bit as well;
  32-bit float ("F")  double accumulate in window order, store as float32 after each pass.
Horizontal pass first, then vertical; a pass whose size does not change is skipped.

Pinned bit-exactly against Pillow itself in tests/test_resize_oracle.py.
"""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def bicubic(x, a=-0.5):
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def coeffs(in_size, out_size):
    """-> (bounds int32 [out, 2] = (xmin, count), weights float64 [out, ksize], ksize)"""
    scale = filterscale = in_size / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.float64)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        n = xmax - xmin
        ww = 0.0
        for x in range(n):
            w = bicubic((x + xmin - center + 0.5) * ss)
            kk[xx, x] = w
            ww += w
        if ww != 0.0:
            for x in range(n):
                kk[xx, x] /= ww
        bounds[xx] = (xmin, n)
    return bounds, kk, ksize


def fixed_point(kk):
    """normalize_coeffs_8bpc: (int)(+-0.5 + k * 2^22), truncation toward zero."""
    s = kk * float(1 << PRECISION_BITS)
    return np.where(kk < 0, np.trunc(-0.5 + s), np.trunc(0.5 + s)).astype(np.int32)


def _pass_u8(img, bounds, ik, axis):
    """img uint8 [H, W, C]; resample along `axis` (1 = horizontal, 0 = vertical)."""
    src = np.moveaxis(img, axis, 0).astype(np.int64)               # [in, other, C]
    out = np.empty((bounds.shape[0],) + src.shape[1:], dtype=np.uint8)
    for i, (lo, n) in enumerate(bounds):
        acc = np.full(src.shape[1:], 1 << (PRECISION_BITS - 1), dtype=np.int64)
        for x in range(n):
            acc += src[lo + x] * int(ik[i, x])
        acc = ((acc + (1 << 31)) & 0xFFFFFFFF) - (1 << 31)         # int32 wrap-around of the C accumulator (never hit)
        out[i] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def _pass_f32(img, bounds, kk, axis):
    src = np.moveaxis(img, axis, 0).astype(np.float64)
    out = np.empty((bounds.shape[0],) + src.shape[1:], dtype=np.float32)
    for i, (lo, n) in enumerate(bounds):
        acc = np.zeros(src.shape[1:], dtype=np.float64)
        for x in range(n):
            acc = acc + src[lo + x] * kk[i, x]
        out[i] = acc.astype(np.float32)
    return np.moveaxis(out, 0, axis)


def resize_u8(img, out_w, out_h):
    """img uint8 [H, W, C] -> uint8 [out_h, out_w, C], equal to np.array(Image.fromarray(img).resize((out_w, out_h)))."""
    H, W = img.shape[:2]
    if W != out_w:
        b, k, _ = coeffs(W, out_w)
        img = _pass_u8(img, b, fixed_point(k), 1)
    if H != out_h:
        b, k, _ = coeffs(H, out_h)
        img = _pass_u8(img, b, fixed_point(k), 0)
    return img


def resize_f32(img, out_w, out_h):
    """img float32 [H, W] -> float32 [out_h, out_w], equal to the same call on a mode-"F" image."""
    H, W = img.shape
    if W != out_w:
        b, k, _ = coeffs(W, out_w)
        img = _pass_f32(img, b, k, 1)
    if H != out_h:
        b, k, _ = coeffs(H, out_h)
        img = _pass_f32(img, b, k, 0)
    return img
