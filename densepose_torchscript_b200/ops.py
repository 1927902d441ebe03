"""Per-op Python entry points over the C-ABI (used by the stage-parity tests and tools).

torch is plumbing here: it owns device memory and the current stream; every op below enqueues
hand-written sm_100a kernels from libdpb200.so and nothing else.
"""
import ctypes as C
import math
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import lib, check


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def round_up(v: int, m: int) -> int:
    return (v + m - 1) // m * m


def split_bf16(x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """fp32 -> the strict mode's bf16 pair (hi, lo): hi = bf16(x), lo = bf16(x - hi); hi + lo carries 16 mantissa bits."""
    hi = x.float().to(torch.bfloat16)
    return hi, (x.float() - hi.float()).to(torch.bfloat16)


def pack_conv_weight(w: torch.Tensor, bias: Optional[torch.Tensor] = None, strict: bool = False):
    """OIHW fp32 conv weight (or [out,in] linear weight) -> K-major bf16 [cout_pad, kh*kw*cin_pad] (strict: three K
    segments [w_hi | w_lo | w_hi] per tap). Returns (packed, bias_fp32[cout_pad], cin_pad, cout_pad)."""
    if w.dim() == 2:
        w = w[:, :, None, None]
    co, ci, kh, kw = w.shape
    cin_pad, cout_pad = round_up(ci, 64), round_up(co, 16)
    p = torch.zeros(cout_pad, kh, kw, cin_pad, dtype=torch.float32, device=w.device)
    p[:co, :, :, :ci] = w.detach().float().permute(0, 2, 3, 1)
    if strict:
        hi, lo = split_bf16(p)
        packed = torch.stack([hi, lo, hi], dim=3).reshape(cout_pad, kh * kw * 3 * cin_pad).contiguous()
    else:
        packed = p.reshape(cout_pad, kh * kw * cin_pad).to(torch.bfloat16).contiguous()
    b = torch.zeros(cout_pad, dtype=torch.float32, device=w.device)
    if bias is not None:
        b[:co] = bias.detach().float()
    return packed, b, cin_pad, cout_pad


def conv2d(x: torch.Tensor, packed: torch.Tensor, bias: Optional[torch.Tensor], kh: int, kw: int,
           stride: int = 1, pad: int = 0, dil: int = 1, relu: bool = False,
           res: Optional[torch.Tensor] = None, res_shift: int = 0, out_fp32: bool = False,
           n_valid: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
           block_n: int = 0, stages: int = 0, tiled: bool = False, epilogue: int = 0,
           planar: bool = False, ks: int = 0, phase_taps: bool = False, pair: int = 0,
           x_lo: Optional[torch.Tensor] = None, res_lo: Optional[torch.Tensor] = None,
           out_lo: Optional[torch.Tensor] = None):
    """x: bf16 NHWC [N,H,W,Cin] (contiguous). Returns NHWC [N,Ho,Wo,cout_pad] bf16 (or fp32); with
    planar=True the fp32 result is channel-planar [N,cout_pad,Ho,Wo].
    Strict mode (x_lo given): x / res are bf16 (hi, lo) pairs, `packed` comes from pack_conv_weight(strict=True); a bf16
    result is returned as the pair (hi, lo)."""
    _lib.require_device()
    assert x.is_cuda and x.dtype == torch.bfloat16 and x.is_contiguous() and x.dim() == 4
    n, h, w, cin = x.shape
    cout_pad, ktot = packed.shape
    strict = x_lo is not None
    cin_pad = ktot // (kh * kw * (3 if strict else 1))
    ho = (h + 2 * pad - dil * (kh - 1) - 1) // stride + 1
    wo = (w + 2 * pad - dil * (kw - 1) - 1) // stride + 1
    if phase_taps:          # ConvTranspose2d(k4,s2,p1) phases: one low-res output pixel per input pixel
        ho, wo = h, w
    if out is None:
        shape = (n, cout_pad, ho, wo) if planar else (n, ho, wo, cout_pad)
        out = torch.empty(*shape, device=x.device, dtype=torch.float32 if (out_fp32 or planar) else torch.bfloat16)
    a = _lib.Conv2dArgs()
    a.x = x.data_ptr(); a.n, a.h, a.w, a.cin = n, h, w, cin
    a.x_sn = a.x_sh = a.x_sw = 0
    a.wgt = packed.data_ptr(); a.cin_pad, a.cout_pad = cin_pad, cout_pad
    a.bias = _p(bias)
    a.kh, a.kw, a.sy, a.sx, a.pad_y, a.pad_x, a.dil = kh, kw, stride, stride, pad, pad, dil
    a.h_out, a.w_out, a.relu = ho, wo, int(relu)
    if res is not None:
        assert res.dtype == torch.bfloat16 and res.is_contiguous()
        a.res = res.data_ptr()
        a.res_sx = res.shape[3]; a.res_sy = res.shape[3] * res.shape[2]
        a.res_sn = res.shape[3] * res.shape[2] * res.shape[1]
    a.res_shift = res_shift
    a.y = out.data_ptr(); a.y_fp32 = int(out.dtype == torch.float32)
    if planar:
        a.y_sx = out.stride(3); a.y_sy = out.stride(2); a.y_sn = out.stride(0); a.y_sc = out.stride(1)
    else:
        a.y_sx = out.stride(2); a.y_sy = out.stride(1); a.y_sn = out.stride(0); a.y_sc = 1
    a.epilogue = epilogue
    a.n_valid = _p(n_valid)
    a.block_n, a.stages, a.tiled, a.ks = block_n, stages, int(tiled), ks
    a.phase_taps = int(phase_taps)
    a.pair = pair
    if strict:
        assert x_lo.dtype == torch.bfloat16 and x_lo.shape == x.shape and x_lo.is_contiguous()
        a.x_lo = x_lo.data_ptr()
        if res is not None:
            assert res_lo is not None and res_lo.shape == res.shape and res_lo.is_contiguous()
            a.res_lo = res_lo.data_ptr()
        if out.dtype == torch.bfloat16:
            out_lo = torch.empty_like(out) if out_lo is None else out_lo
            a.y_lo = out_lo.data_ptr()
    check(lib.dpb200_conv2d(C.byref(a), _stream()), "dpb200_conv2d")
    return (out, out_lo) if (strict and out.dtype == torch.bfloat16) else out


def preprocess(images: torch.Tensor, k: float, mean: Sequence[float], std: Sequence[float],
               flip_rgb: bool = False, variant: int = 0) -> Tuple[torch.Tensor, Tuple[int, int, int, int]]:
    """images [B,H0,W0,3] fp32/u8 cuda -> space-to-depth stem layout [B,Hp/2,Wp/2+4,16] bf16 (see `stem_to_image`).
    uint8 images go through ATen's fixed-point uint8 bilinear (tables built on the device first); `variant` picks the
    float kernel of a multi-threaded (0) or single-threaded (1) reference. Returns (dst, (Hr,Wr,Hp,Wp))."""
    _lib.require_device()
    b, h0, w0, _ = images.shape
    hr, wr = int(math.floor(h0 * k)), int(math.floor(w0 * k))
    hp, wp = round_up(hr, 32), round_up(wr, 32)
    dst = torch.empty(b, hp // 2, wp // 2 + 4, 16, dtype=torch.bfloat16, device=images.device)
    a = _lib.PreprocessArgs()
    a.src = images.data_ptr(); a.src_u8 = int(images.dtype == torch.uint8)
    a.b, a.h0, a.w0, a.hr, a.wr = b, h0, w0, hr, wr
    a.inv_scale = float(torch.tensor(1.0 / k, dtype=torch.float64).to(torch.float32))
    a.flip_rgb = int(flip_rgb)
    a.variant = variant
    for i in range(3):
        a.mean[i] = mean[i]; a.std[i] = std[i]
    a.dst = dst.data_ptr(); a.hp, a.wx = hp, wp // 2 + 4
    if images.dtype == torch.uint8:
        tables = torch.empty(1 + hr + wr, 2, dtype=torch.int32, device=images.device)
        check(lib.dpb200_u8_resize_tables(tables.data_ptr(), h0, hr, w0, wr, 1.0 / k, _stream()), "dpb200_u8_resize_tables")
        a.tables = tables.data_ptr()
    check(lib.dpb200_preprocess(C.byref(a), _stream()), "dpb200_preprocess")
    return dst, (hr, wr, hp, wp)


def stem_to_image(dst: torch.Tensor) -> torch.Tensor:
    """Inverse of the stem layout: [B,Hp/2,Wq,16] -> [B,Hp,2*Wq,4] (padded-image pixel (y,x) at column x + 4)."""
    b, hq, wq, _ = dst.shape
    return dst.view(b, hq, wq, 2, 2, 4).permute(0, 1, 3, 2, 4, 5).reshape(b, 2 * hq, 2 * wq, 4)


def maxpool3x3s2(x: torch.Tensor) -> torch.Tensor:
    b, h, w, c = x.shape
    y = torch.empty(b, (h - 1) // 2 + 1, (w - 1) // 2 + 1, c, dtype=torch.bfloat16, device=x.device)
    check(lib.dpb200_maxpool3x3s2(x.data_ptr(), y.data_ptr(), b, h, w, c, _stream()), "dpb200_maxpool3x3s2")
    return y


def upsample2x(x: torch.Tensor) -> torch.Tensor:
    b, h, w, c = x.shape
    y = torch.empty(b, 2 * h, 2 * w, c, dtype=torch.bfloat16, device=x.device)
    check(lib.dpb200_upsample2x(x.data_ptr(), y.data_ptr(), b, h, w, c, _stream()), "dpb200_upsample2x")
    return y


def decoder_merge(a: torch.Tensor, b3: torch.Tensor, b4: torch.Tensor, b5: torch.Tensor) -> torch.Tensor:
    b, h, w, c = a.shape
    out = torch.empty_like(a)
    check(lib.dpb200_decoder_merge(a.data_ptr(), b3.data_ptr(), b4.data_ptr(), b5.data_ptr(), out.data_ptr(),
                                   b, h, w, c, _stream()), "dpb200_decoder_merge")
    return out


def cell_anchors(size: float) -> List[float]:
    out = []
    for ar in (0.5, 1.0, 2.0):
        w = math.sqrt(size * size / ar)
        h = ar * w
        out += [-w / 2.0, -h / 2.0, w / 2.0, h / 2.0]
    return out


def rpn_proposals(heads: List[torch.Tensor], clip_x: float, clip_y: float, pre_topk: int = 1000,
                  post_topk: int = 1000, nms_thresh: float = 0.7):
    """heads[l]: [B,H,W,16] fp32 (3 logits, 12 deltas, pad). Returns (boxes [B,post,4], scores, counts, debug)."""
    _lib.require_device()
    b = heads[0].shape[0]
    dev = heads[0].device
    a = _lib.RpnArgs()
    for l, h in enumerate(heads):
        assert h.dtype == torch.float32 and h.is_contiguous() and h.shape[3] == 16
        a.head[l] = h.data_ptr(); a.h[l] = h.shape[1]; a.w[l] = h.shape[2]; a.stride[l] = float(4 << l)
        ca = cell_anchors(float(32 << l))
        for i in range(12):
            a.anchors[l][i] = ca[i]
    a.b, a.pre_topk, a.post_topk, a.nms_thresh, a.clip_x, a.clip_y = b, pre_topk, post_topk, nms_thresh, clip_x, clip_y
    cand_boxes = torch.zeros(b, 5, pre_topk, 4, device=dev)
    cand_scores = torch.zeros(b, 5, pre_topk, device=dev)
    cand_count = torch.zeros(b, 5, dtype=torch.int32, device=dev)
    cand_keep = torch.zeros(b, 5, pre_topk, dtype=torch.uint8, device=dev)
    boxes = torch.zeros(b, post_topk, 4, device=dev)
    scores = torch.zeros(b, post_topk, device=dev)
    counts = torch.zeros(b, dtype=torch.int32, device=dev)
    a.cand_boxes, a.cand_scores, a.cand_count, a.cand_keep = (cand_boxes.data_ptr(), cand_scores.data_ptr(),
                                                               cand_count.data_ptr(), cand_keep.data_ptr())
    a.prop_boxes, a.prop_scores, a.prop_count = boxes.data_ptr(), scores.data_ptr(), counts.data_ptr()
    check(lib.dpb200_rpn_proposals(C.byref(a), _stream()), "dpb200_rpn_proposals")
    return boxes, scores, counts, dict(cand_boxes=cand_boxes, cand_scores=cand_scores, cand_count=cand_count,
                                       cand_keep=cand_keep)


def nms_sorted(boxes: torch.Tensor, thr: float) -> torch.Tensor:
    """boxes [n,4] fp32 sorted by descending score -> bool keep mask [n]."""
    n = boxes.shape[0]
    keep = torch.ones(max(n, 1), dtype=torch.uint8, device=boxes.device)
    if n:
        check(lib.dpb200_nms_sorted(boxes.contiguous().data_ptr(), n, thr, keep.data_ptr(), _stream()),
              "dpb200_nms_sorted")
    return keep[:n].bool()


def roi_align(feats: List[torch.Tensor], rois: torch.Tensor, out_size: int, scales: Sequence[float],
              out_fp32: bool = False, n_rois: Optional[torch.Tensor] = None) -> torch.Tensor:
    """feats[l] NHWC bf16 [B,H,W,C]; rois [R,5] fp32 -> [R,P,P,C]."""
    _lib.require_device()
    a = _lib.RoiAlignArgs()
    for l, f in enumerate(feats):
        assert f.dtype == torch.bfloat16 and f.is_contiguous()
        a.feat[l] = f.data_ptr(); a.h[l] = f.shape[1]; a.w[l] = f.shape[2]; a.scale[l] = scales[l]
    c = feats[0].shape[3]
    r = rois.shape[0]
    a.n_levels, a.c, a.rois, a.n_rois, a.r, a.p = len(feats), c, rois.contiguous().data_ptr(), _p(n_rois), r, out_size
    out = torch.zeros(r, out_size, out_size, c, device=rois.device, dtype=torch.float32 if out_fp32 else torch.bfloat16)
    a.out, a.out_fp32 = out.data_ptr(), int(out_fp32)
    check(lib.dpb200_roi_align(C.byref(a), _stream()), "dpb200_roi_align")
    return out


def box_predict(head: torch.Tensor, prop_boxes: torch.Tensor, prop_count: torch.Tensor, score_thresh: float,
                nms_thresh: float, topk: int, scale_x: float, scale_y: float, out_w: float, out_h: float):
    """head [B*R,16] fp32; prop_boxes [B,R,4]; prop_count [B] int32."""
    _lib.require_device()
    b, r = prop_boxes.shape[0], prop_boxes.shape[1]
    dev = head.device
    a = _lib.BoxPredictArgs()
    a.head, a.prop_boxes, a.prop_count, a.b, a.r = head.data_ptr(), prop_boxes.data_ptr(), prop_count.data_ptr(), b, r
    a.score_thresh, a.nms_thresh, a.topk = score_thresh, nms_thresh, topk
    a.scale_x, a.scale_y, a.out_w, a.out_h = scale_x, scale_y, out_w, out_h
    ws_boxes = torch.zeros(b, 1024, 4, device=dev); ws_keep = torch.zeros(b, 1024, dtype=torch.uint8, device=dev)
    raw = torch.zeros(b, topk, 4, device=dev); boxes = torch.zeros(b, topk, 4, device=dev)
    scores = torch.zeros(b, topk, device=dev); count = torch.zeros(b, dtype=torch.int32, device=dev)
    a.ws_boxes, a.ws_keep = ws_boxes.data_ptr(), ws_keep.data_ptr()
    a.det_boxes_raw, a.det_boxes, a.det_scores, a.det_count = raw.data_ptr(), boxes.data_ptr(), scores.data_ptr(), count.data_ptr()
    check(lib.dpb200_box_predict(C.byref(a), _stream()), "dpb200_box_predict")
    return raw, boxes, scores, count


def groupnorm_relu(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, out_hw: Optional[int] = None,
                   out: Optional[torch.Tensor] = None, n_valid: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x [R,HW,C] bf16 -> [R,out_hw,C] bf16. `out` may be a channel slice of a wider [R,out_hw,Ctot] tensor (the ASPP
    concat, deeplab.py:141); `n_valid` a device int32 count of the ROIs actually present."""
    r, hw, c = x.shape
    out_hw = hw if out_hw is None else out_hw
    y = torch.empty(r, out_hw, c, dtype=torch.bfloat16, device=x.device) if out is None else out
    assert y.dtype == torch.bfloat16 and y.stride(2) == 1 and y.stride(0) == out_hw * y.stride(1)
    check(lib.dpb200_groupnorm_relu(x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), y.data_ptr(), r, hw, c, y.stride(1),
                                    out_hw, _p(n_valid), _stream()), "dpb200_groupnorm_relu")
    return y


def avgpool(x: torch.Tensor) -> torch.Tensor:
    r, hw, c = x.shape
    y = torch.empty(r, c, dtype=torch.bfloat16, device=x.device)
    check(lib.dpb200_avgpool(x.data_ptr(), y.data_ptr(), r, hw, c, None, _stream()), "dpb200_avgpool")
    return y


def predictor_upsample(low: torch.Tensor, channels: Sequence[int]):
    """low: the phase-planar fp32 [R,2,2,Cpad,S/2,S/2] the deconv GEMM writes (low-res pixel (2*yy+py, 2*xx+px) at
    [r,py,px,c,yy,xx]); `channels`: how its leading channels split into heads, e.g. (kc, 25, 25, 25) for coarse / fine /
    u / v (+ the confidence heads of a WC* model). Returns one NCHW fp32 [R,ch,2S,2S] tensor per head."""
    r, _, _, cpad, sh, _ = low.shape
    s = 2 * sh
    dev = low.device
    assert low.is_contiguous() and low.dtype == torch.float32
    outs = [torch.empty(r, c, 2 * s, 2 * s, device=dev) for c in channels]
    ptrs = (C.c_void_p * len(outs))(*[o.data_ptr() for o in outs])
    ch = (C.c_int32 * len(outs))(*channels)
    check(lib.dpb200_predictor_upsample(low.data_ptr(), r, s, cpad, None, ptrs, ch, len(outs), _stream()),
          "dpb200_predictor_upsample")
    return outs


def box_sizes(boxes_xyxy: torch.Tensor):
    """Host-side geometry of the extractor (visualizer.py:40-43,20-30): xywh, `.long()` truncation, max(int, 1).
    Returns (boxes_xywh, wh int32 [D,2] on the CPU, pixel offsets int64 [D+1] on the CPU)."""
    boxes_xywh = boxes_xyxy.clone()
    boxes_xywh[:, 2:] -= boxes_xywh[:, :2]                       # visualizer.py:41-42
    wh = boxes_xywh[:, 2:].long().clamp(min=1).to(torch.int32).cpu()   # .long() truncation, max(int, 1)
    sizes = (wh[:, 0].long() * wh[:, 1].long())
    offsets = torch.zeros(boxes_xyxy.shape[0] + 1, dtype=torch.int64)
    offsets[1:] = torch.cumsum(sizes, 0)
    return boxes_xywh, wh, offsets


def dp_resample_into(coarse, fine, u, v, wh_dev: torch.Tensor, off_dev: torch.Tensor, total: int,
                     labels: torch.Tensor, uv: torch.Tensor, stream=None):
    """Launches the extractor kernel into caller-owned packed buffers (labels int64 or uint8, uv fp32)."""
    a = _lib.ResampleArgs()
    a.coarse, a.fine, a.u, a.v = coarse.data_ptr(), fine.data_ptr(), u.data_ptr(), v.data_ptr()
    a.d, a.kc, a.s = wh_dev.shape[0], coarse.shape[1], coarse.shape[2]
    a.box_wh, a.offsets, a.labels, a.uv, a.total_pixels = (wh_dev.data_ptr(), off_dev.data_ptr(), labels.data_ptr(),
                                                           uv.data_ptr(), total)
    a.labels_u8 = int(labels.dtype == torch.uint8)
    check(lib.dpb200_dp_resample(C.byref(a), stream if stream is not None else _stream()), "dpb200_dp_resample")


def dp_resample(coarse: torch.Tensor, fine: torch.Tensor, u: torch.Tensor, v: torch.Tensor,
                boxes_xyxy: torch.Tensor, labels_u8: bool = False):
    """DensePoseResultExtractor on the device. Returns (list of {'labels','uv'} per box, boxes_xywh)."""
    _lib.require_device()
    d = boxes_xyxy.shape[0]
    dev = coarse.device
    boxes_xywh, wh, offsets = box_sizes(boxes_xyxy)
    total = int(offsets[-1])
    labels = torch.empty(max(total, 1), dtype=torch.uint8 if labels_u8 else torch.int64, device=dev)
    uv = torch.empty(max(2 * total, 1), dtype=torch.float32, device=dev)
    if d and total:
        dp_resample_into(coarse.contiguous(), fine.contiguous(), u.contiguous(), v.contiguous(),
                         wh.to(dev).contiguous(), offsets.to(dev), total, labels, uv)
    results = []
    for i in range(d):
        w, h = int(wh[i, 0]), int(wh[i, 1])
        o = int(offsets[i])
        results.append({"labels": labels[o:o + h * w].view(h, w), "uv": uv[2 * o:2 * o + 2 * h * w].view(2, h, w)})
    return results, boxes_xywh
