"""Per-op Python entry points over the C-ABI (used by the stage-parity tests and the weight packer).

torch is plumbing here: it owns device memory and the current stream; every op below enqueues
hand-written sm_100a kernels from libdpb200.so and nothing else.
"""
import ctypes as C
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import lib, check


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def round_up(v: int, m: int) -> int:
    return (v + m - 1) // m * m


def pack_conv_weight(w: torch.Tensor, bias: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor, int, int]:
    """OIHW fp32 conv weight (or [out,in] linear weight) -> K-major bf16 [cout_pad, kh*kw*cin_pad].

    K index = (ky*kw + kx)*cin_pad + ci, matching the NHWC slab order the TMA producer walks.
    Returns (packed, bias_fp32[cout_pad], cin_pad, cout_pad).
    """
    if w.dim() == 2:
        w = w[:, :, None, None]
    co, ci, kh, kw = w.shape
    cin_pad, cout_pad = round_up(ci, 64), round_up(co, 16)
    p = torch.zeros(cout_pad, kh, kw, cin_pad, dtype=torch.float32, device=w.device)
    p[:co, :, :, :ci] = w.detach().float().permute(0, 2, 3, 1)
    packed = p.reshape(cout_pad, kh * kw * cin_pad).to(torch.bfloat16).contiguous()
    b = torch.zeros(cout_pad, dtype=torch.float32, device=w.device)
    if bias is not None:
        b[:co] = bias.detach().float()
    return packed, b, cin_pad, cout_pad


def conv2d(x: torch.Tensor, packed: torch.Tensor, bias: Optional[torch.Tensor], kh: int, kw: int,
           stride: int = 1, pad: int = 0, dil: int = 1, relu: bool = False,
           res: Optional[torch.Tensor] = None, res_shift: int = 0, out_fp32: bool = False,
           n_valid: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
           block_n: int = 0, stages: int = 0, tiled: bool = False, pad_xy: Optional[Tuple[int, int]] = None) -> torch.Tensor:
    """x: bf16 NHWC [N,H,W,Cin] (contiguous). Returns NHWC [N,Ho,Wo,cout_pad] bf16 (or fp32)."""
    _lib.require_device()
    assert x.is_cuda and x.dtype == torch.bfloat16 and x.is_contiguous() and x.dim() == 4
    n, h, w, cin = x.shape
    cout_pad, ktot = packed.shape
    cin_pad = ktot // (kh * kw)
    pad_y, pad_x = (pad, pad) if pad_xy is None else pad_xy
    ho = (h + 2 * pad_y - dil * (kh - 1) - 1) // stride + 1 if pad_xy is None else h
    wo = (w + 2 * pad_x - dil * (kw - 1) - 1) // stride + 1 if pad_xy is None else w
    if out is None:
        out = torch.empty(n, ho, wo, cout_pad, device=x.device,
                          dtype=torch.float32 if out_fp32 else torch.bfloat16)
    a = _lib.Conv2dArgs()
    a.x = x.data_ptr(); a.n, a.h, a.w, a.cin = n, h, w, cin
    a.x_sn = a.x_sh = a.x_sw = 0
    a.wgt = packed.data_ptr(); a.cin_pad, a.cout_pad = cin_pad, cout_pad
    a.bias = bias.data_ptr() if bias is not None else None
    a.kh, a.kw, a.sy, a.sx, a.pad_y, a.pad_x, a.dil = kh, kw, stride, stride, pad_y, pad_x, dil
    a.h_out, a.w_out, a.relu = ho, wo, int(relu)
    if res is not None:
        assert res.dtype == torch.bfloat16 and res.is_contiguous()
        a.res = res.data_ptr()
        a.res_sx = res.shape[3]; a.res_sy = res.shape[3] * res.shape[2]
        a.res_sn = res.shape[3] * res.shape[2] * res.shape[1]
    a.res_shift = res_shift
    a.y = out.data_ptr(); a.y_fp32 = int(out.dtype == torch.float32)
    a.y_sx = out.stride(2); a.y_sy = out.stride(1); a.y_sn = out.stride(0)
    a.n_valid = n_valid.data_ptr() if n_valid is not None else None
    a.block_n, a.stages, a.tiled = block_n, stages, int(tiled)
    check(lib.dpb200_conv2d(C.byref(a), _stream()), "dpb200_conv2d")
    return out
