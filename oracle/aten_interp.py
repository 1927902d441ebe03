"""CPU ORACLE — TEST INFRASTRUCTURE ONLY (never imported by the product path).

Restatement of the ATen CPU `upsample_bilinear2d` kernels the reference reaches through `F.interpolate`
(defaults.py:89 image resize; visualizer.py:14-25 per-box resample; chart.py:72-74 predictor tail). ATen is a
third-party dependency of the reference (requirements.txt:1 pins torch~=2.1.2, this image has 2.11.0) whose source is
not under /root/reference, so the arithmetic below was probed on the installed build and is pinned against it in
tests/test_oracle.py (`test_aten_*`): every function here reproduces `F.interpolate` BIT FOR BIT.

What the probe found (aten/src/ATen/native/cpu/UpSampleKernel.cpp, UpSampleKernelAVXAntialias.h; x86 AVX2 / AVX512
builds, both with 8-lane float vectors):

uint8 (what run.py:33-36 feeds): Pillow-style fixed point, horizontal pass then vertical pass, each rounding to uint8.
  Per axis: scale = 1.0 / k (double); for output i: center = scale * (i + 0.5); first tap xmin = max(int(center - 1 +
  0.5), 0), taps = min(int(center + 1 + 0.5), in) - xmin (at most 2); weights w_j = max(0, 1 - |j + xmin - center +
  0.5|), normalised by their sum; precision p = the first p in [0, 22) with int(0.5 + w_max * 2^(p+1)) >= 2^15;
  int16 weights = int(0.5 + w * 2^p); pass = clip8((sum_j w_j * pixel_j + 2^(p-1)) >> p).

float: source index real = fma(scale, i + 0.5, -0.5) clamped at 0 (ONE fused multiply-add), i0 = min(int(real), in-1),
  i1 = i0 + (i0 < in-1), l1 = clamp(real - i0, 0, 1), l0 = 1 - l1. Then one of two kernels:
  * separable generic kernel (contiguous NCHW input, output h + w > 128; also the channels-last-strided 3-channel image
    of defaults.py:89 when ATen runs with more than one intra-op thread):
        top = fma(lx0, p00, lx1 * p01); bot = fma(lx0, p10, lx1 * p11); out = fma(ly0, top, ly1 * bot)
  * channels-last kernel (output h + w <= 128; channels-last input with C > 3; or the 3-channel image with ONE
    thread): w_ij = ly_i * lx_j (rounded); channels below C - C % 8 use the vector expression
        s = fma(w11, p11, w10 * p10); s = fma(w01, p01, s); s = fma(w00, p00, s)
    and the remaining (tail) channels the scalar one
        s = fma(w00, p00, w01 * p01); s = fma(w10, p10, s); s = fma(w11, p11, s)
  So the reference's own fp32 results depend on the output size, the channel index and the thread count; the CUDA
  kernels (elementwise.cu, resample.cu) take the same branches.
"""
import math
from typing import Tuple

import numpy as np

f32 = np.float32


def fma32(a, b, c):
    """fp32 fused multiply-add. a*b is exact in a 64-bit mantissa (24 + 24 bits); the sum is rounded to 64 bits and then
    to 24 — a double rounding only when the 64-bit value lands exactly on a 24-bit tie, which needs > 40 cancelling
    bits and does not occur in practice (the pin test compares ~10^7 values)."""
    return (np.asarray(a, np.longdouble) * np.asarray(b, np.longdouble) + np.asarray(c, np.longdouble)).astype(np.float32)


def mul32(a, b):
    return (np.asarray(a, f32) * np.asarray(b, f32)).astype(np.float32)


# ------------------------------------------------------------------------------------------------ uint8
def u8_axis_table(in_size: int, out_size: int, scale: float) -> Tuple[np.ndarray, np.ndarray, int]:
    """(first tap index [out], int16 weights [out, 2], precision) of one axis; `scale` = 1.0 / k in double."""
    xmin = np.zeros(out_size, np.int64)
    w = np.zeros((out_size, 2), np.float64)
    for i in range(out_size):
        center = scale * (i + 0.5)
        lo = max(int(center - 1.0 + 0.5), 0)
        n = max(0, min(min(int(center + 1.0 + 0.5), in_size) - lo, 2))
        ws = [max(0.0, 1.0 - abs(j + lo - center + 0.5)) for j in range(n)]
        tot = 0.0
        for v in ws:
            tot += v
        for j in range(n):
            w[i, j] = ws[j] / tot if tot != 0.0 else ws[j]
        xmin[i] = lo
    wmax = float(w.max()) if out_size else 0.0
    p = 0
    while p < 22:
        if int(0.5 + wmax * (1 << (p + 1))) >= (1 << 15):
            break
        p += 1
    wi = np.floor(0.5 + w * float(1 << p)).astype(np.int64)
    return xmin, wi, p


def upsample_bilinear_u8(img_hwc: np.ndarray, k: float) -> np.ndarray:
    """F.interpolate(uint8 [1,3,H,W], scale_factor=k, mode='bilinear', align_corners=False) -> HWC uint8."""
    assert img_hwc.dtype == np.uint8
    H, W = img_hwc.shape[:2]
    Ho, Wo = int(math.floor(H * k)), int(math.floor(W * k))
    scale = 1.0 / k
    xm, xw, px = u8_axis_table(W, Wo, scale)
    ym, yw, py = u8_axis_table(H, Ho, scale)
    a = img_hwc.astype(np.int64)
    x1 = np.minimum(xm + 1, W - 1)
    tmp = (a[:, xm] * xw[None, :, 0, None] + a[:, x1] * xw[None, :, 1, None] + (1 << (px - 1))) >> px
    tmp = np.clip(tmp, 0, 255)
    y1 = np.minimum(ym + 1, H - 1)
    out = (tmp[ym] * yw[:, 0, None, None] + tmp[y1] * yw[:, 1, None, None] + (1 << (py - 1))) >> py
    return np.clip(out, 0, 255).astype(np.uint8)


# ------------------------------------------------------------------------------------------------ float
def f32_axis(scale: np.float32, n_out: int, n_in: int):
    d = np.arange(n_out, dtype=np.float32)
    real = np.maximum(fma32(f32(scale), d + f32(0.5), f32(-0.5)), f32(0)).astype(np.float32)
    i0 = np.minimum(real.astype(np.int64), n_in - 1)
    i1 = np.minimum(i0 + 1, n_in - 1)
    l1 = np.clip((real - i0.astype(np.float32)).astype(np.float32), 0, 1).astype(np.float32)
    return i0, i1, (f32(1) - l1).astype(np.float32), l1


def upsample_bilinear_f32(x: np.ndarray, out_h: int, out_w: int, scale_h=None, scale_w=None,
                          kernel: str = "auto") -> np.ndarray:
    """F.interpolate(float32 [N,C,H,W], ..., mode='bilinear', align_corners=False). `scale_*` = the scale_factor given
    to F.interpolate (None: size given). kernel: 'separable' | 'channels_last' | 'auto' (ATen's choice for a
    contiguous NCHW input: channels_last iff out_h + out_w <= 128)."""
    assert x.dtype == np.float32 and x.ndim == 4
    N, C, H, W = x.shape
    sh = f32(1.0 / scale_h) if scale_h else f32(H) / f32(out_h)
    sw = f32(1.0 / scale_w) if scale_w else f32(W) / f32(out_w)
    y0, y1, ly0, ly1 = f32_axis(sh, out_h, H)
    x0, x1, lx0, lx1 = f32_axis(sw, out_w, W)
    p00, p01 = x[:, :, y0][:, :, :, x0], x[:, :, y0][:, :, :, x1]
    p10, p11 = x[:, :, y1][:, :, :, x0], x[:, :, y1][:, :, :, x1]
    LX0, LX1 = lx0[None, None, None, :], lx1[None, None, None, :]
    LY0, LY1 = ly0[None, None, :, None], ly1[None, None, :, None]
    if kernel == "auto":
        kernel = "channels_last" if out_h + out_w <= 128 else "separable"
    if kernel == "separable":
        top = fma32(LX0, p00, mul32(LX1, p01))
        bot = fma32(LX0, p10, mul32(LX1, p11))
        return fma32(LY0, top, mul32(LY1, bot))
    w00, w01, w10, w11 = mul32(LY0, LX0), mul32(LY0, LX1), mul32(LY1, LX0), mul32(LY1, LX1)
    nv = C - C % 8
    out = np.empty((N, C, out_h, out_w), np.float32)
    if nv:
        s = fma32(w11, p11[:, :nv], mul32(w10, p10[:, :nv]))
        s = fma32(w01, p01[:, :nv], s)
        out[:, :nv] = fma32(w00, p00[:, :nv], s)
    if nv < C:
        s = fma32(w00, p00[:, nv:], mul32(w01, p01[:, nv:]))
        s = fma32(w10, p10[:, nv:], s)
        out[:, nv:] = fma32(w11, p11[:, nv:], s)
    return out
