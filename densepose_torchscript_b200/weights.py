"""Load-time weight pre-pack: reference state_dict -> the engine's named, GEMM-ready parameters.

Accepts the reference's key families and either alias of each duplicated registration (SURVEY.md §8
quirk 10: res{2..5}|stages.N, fpn_lateral/outputN|lateral/output_convs.N, decoder.pN|scale_heads.N,
body_conv_fcnN|stacked_convs.N), with or without the predictor's "model." prefix.

Transforms (all exact re-layouts except the bf16 rounding of the final weights):
  * FrozenBatchNorm2d folded into conv weight + fp32 bias (detectron2/layers/batch_norm.py:45-46, eps 1e-5);
  * OIHW -> K-major [cout_pad][(ky*kw+kx)*cin_pad + ci] bf16 (NHWC implicit-GEMM operand);
  * stem 7x7/2 -> 4x4/1 over the 2x2 space-to-depth image: per row tap a window of 4 pixels x (dy, dx, c4) channels;
  * RPN objectness (3) + anchor deltas (12) fused into one 16-row 1x1 head; cls_score (2) + bbox_pred (4) likewise;
  * FC1 columns permuted from (c,y,x) to the NHWC pooled order (y,x,c) (box_head.py:70-71);
  * the four ConvTranspose2d(k=4,s=2,p=1) predictors concatenated on Cout and split into four 2x2 output-parity
    phases (densepose/modeling/predictors/chart.py:45-59);
  * DeepLab ASPP rate-56 branch reduced to its centre tap (rate >= pooled size: every other tap reads padding).
"""
from typing import Dict, Optional, Tuple

import torch

from .config import ModelSpec

BN_EPS = 1e-5
Packed = Tuple[torch.Tensor, Optional[torch.Tensor], int, int]


def round_up(v: int, m: int) -> int:
    return (v + m - 1) // m * m


def canonicalize(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Strip 'model.' and map alias keys onto the canonical names."""
    out: Dict[str, torch.Tensor] = {}
    for k, v in sd.items():
        if k.startswith("model."):
            k = k[len("model."):]
        if k.startswith("backbone.bottom_up.stages."):
            rest = k[len("backbone.bottom_up.stages."):]
            idx, tail = rest.split(".", 1)
            k = f"backbone.bottom_up.res{int(idx) + 2}.{tail}"
        for kind in ("lateral", "output"):
            pre = f"backbone.{kind}_convs."
            if k.startswith(pre):
                idx, tail = k[len(pre):].split(".", 1)
                k = f"backbone.fpn_{kind}{5 - int(idx)}.{tail}"
        if k.startswith("roi_heads.decoder.scale_heads."):
            idx, tail = k[len("roi_heads.decoder.scale_heads."):].split(".", 1)
            k = f"roi_heads.decoder.p{int(idx) + 2}.{tail}"
        if k.startswith("roi_heads.densepose_head.stacked_convs."):
            idx, tail = k[len("roi_heads.densepose_head.stacked_convs."):].split(".", 1)
            k = f"roi_heads.densepose_head.body_conv_fcn{int(idx) + 1}.{tail}"
        out.setdefault(k, v)
    return out


def _fold(sd, prefix) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    w = sd[prefix + ".weight"].detach().float()
    frozen_bn = prefix.startswith("backbone.bottom_up.")      # the only FrozenBatchNorm2d layers (DeepLab's `.norm` is GroupNorm)
    if prefix + ".norm.running_var" in sd or (frozen_bn and prefix + ".norm.weight" in sd):
        # checkpoints without running statistics (Caffe2 AffineChannel blobs) load as mean 0 / var 1
        # (FrozenBatchNorm2d._load_from_state_dict, detectron2/layers/batch_norm.py:75-82)
        gamma = sd[prefix + ".norm.weight"].float()
        var = sd[prefix + ".norm.running_var"].float() if prefix + ".norm.running_var" in sd else torch.ones_like(gamma)
        mean = sd[prefix + ".norm.running_mean"].float() if prefix + ".norm.running_mean" in sd else torch.zeros_like(gamma)
        scale = gamma * (var + BN_EPS).rsqrt()
        bias = sd[prefix + ".norm.bias"].float() - mean * scale
        return w * scale.view(-1, 1, 1, 1), bias
    b = sd.get(prefix + ".bias")
    return w, (b.detach().float() if b is not None else None)


def _pack_khwc(w_khwc: torch.Tensor, bias: Optional[torch.Tensor], device, strict: bool = False) -> Packed:
    """w_khwc: [cout, kh, kw, cin] fp32. strict: every tap carries three K segments [w_hi | w_lo | w_hi] with
    w_hi = bf16(w), w_lo = bf16(w - w_hi); the conv kernel pairs them with the activation halves (x_hi, x_hi, x_lo), so
    the fp32 accumulator receives x_hi*w_hi + x_hi*w_lo + x_lo*w_hi (conv_igemm.cuh, strict mode)."""
    co, kh, kw, ci = w_khwc.shape
    cin_pad, cout_pad = round_up(ci, 64), round_up(co, 16)
    p = torch.zeros(cout_pad, kh, kw, cin_pad, dtype=torch.float32)
    p[:co, :, :, :ci] = w_khwc
    if strict:
        hi = p.to(torch.bfloat16)
        lo = (p - hi.float()).to(torch.bfloat16)
        packed = torch.stack([hi, lo, hi], dim=3).reshape(cout_pad, kh * kw * 3 * cin_pad).contiguous().to(device)
    else:
        packed = p.reshape(cout_pad, kh * kw * cin_pad).to(torch.bfloat16).contiguous().to(device)
    b = None
    if bias is not None:
        b = torch.zeros(cout_pad, dtype=torch.float32)
        b[:co] = bias
        b = b.to(device)
    return packed, b, cin_pad, cout_pad


def pack_state_dict(sd: Dict[str, torch.Tensor], spec: ModelSpec, device, strict: bool = False) -> Dict[str, Packed]:
    """strict: weights for the fp32-class numerics mode (three K segments per tap, see _pack_khwc)."""
    sd = {k: v.detach().cpu() for k, v in canonicalize(sd).items()}
    out: Dict[str, Packed] = {}

    def _pack_conv(w_oihw: torch.Tensor, bias, device) -> Packed:
        if w_oihw.dim() == 2:
            w_oihw = w_oihw[:, :, None, None]
        return _pack_khwc(w_oihw.permute(0, 2, 3, 1), bias, device, strict)

    def conv(prefix, with_bias=True):
        w, b = _fold(sd, prefix)
        out[prefix] = _pack_conv(w, b if with_bias else None, device)

    # stem: [64,3,7,7] stride 2 pad 3 -> 4x4 stride 1 over the 2x2 space-to-depth input:
    # [64][ky'][kx'][(dy, dx, c4)] = W[:, c, 2ky'+dy-1, 2kx'+dx-1] (zero outside the 7x7 support)
    w, b = _fold(sd, "backbone.bottom_up.stem.conv1")
    sw = torch.zeros(64, 4, 4, 2, 2, 4)
    for kyq in range(4):
        for dy in range(2):
            ky = 2 * kyq + dy - 1
            if not 0 <= ky <= 6:
                continue
            for kxq in range(4):
                for dx in range(2):
                    kx = 2 * kxq + dx - 1
                    if 0 <= kx <= 6:
                        sw[:, kyq, kxq, dy, dx, :3] = w[:, :, ky, kx]
    out["backbone.bottom_up.stem.conv1"] = _pack_khwc(sw.reshape(64, 4, 1, 64), b, device, strict)
    for si, nb in enumerate(spec.blocks):
        for bi in range(nb):
            p = f"backbone.bottom_up.res{si + 2}.{bi}"
            if bi == 0:
                conv(p + ".shortcut")
            conv(p + ".conv1"); conv(p + ".conv2"); conv(p + ".conv3")
    for lvl in (2, 3, 4, 5):
        conv(f"backbone.fpn_lateral{lvl}"); conv(f"backbone.fpn_output{lvl}")
    rp = "proposal_generator.rpn_head."
    conv(rp + "conv")
    w = torch.cat([sd[rp + "objectness_logits.weight"].float(), sd[rp + "anchor_deltas.weight"].float()], 0)
    b = torch.cat([sd[rp + "objectness_logits.bias"].float(), sd[rp + "anchor_deltas.bias"].float()], 0)
    out[rp + "pred"] = _pack_conv(w, b, device)
    fc1 = sd["roi_heads.box_head.fc1.weight"].float()
    fc1 = fc1.view(fc1.shape[0], 256, 7, 7).permute(0, 2, 3, 1).reshape(fc1.shape[0], -1)
    out["roi_heads.box_head.fc1"] = _pack_conv(fc1, sd["roi_heads.box_head.fc1.bias"].float(), device)
    out["roi_heads.box_head.fc2"] = _pack_conv(sd["roi_heads.box_head.fc2.weight"].float(),
                                               sd["roi_heads.box_head.fc2.bias"].float(), device)
    bp = "roi_heads.box_predictor."
    w = torch.cat([sd[bp + "cls_score.weight"].float(), sd[bp + "bbox_pred.weight"].float()], 0)
    b = torch.cat([sd[bp + "cls_score.bias"].float(), sd[bp + "bbox_pred.bias"].float()], 0)
    out[bp + "pred"] = _pack_conv(w, b, device)
    if spec.decoder_on:
        for name, n in (("p2", 1), ("p3", 1), ("p4", 2), ("p5", 3)):
            for k in range(n):
                conv(f"roi_heads.decoder.{name}.{k if name == 'p2' else 2 * k}")
        conv("roi_heads.decoder.predictor")
    hp = "roi_heads.densepose_head."
    if spec.head == "deeplab":
        def gn(prefix):
            out[prefix] = (sd[prefix + ".weight"].float().contiguous().to(device),
                           sd[prefix + ".bias"].float().contiguous().to(device), 0, 0)
        for i in (0, 1, 2):
            conv(hp + f"ASPP.convs.{i}.0", with_bias=False); gn(hp + f"ASPP.convs.{i}.1")
        if 56 < spec.pooler_res:
            raise ValueError("ASPP rate-56 centre-tap reduction needs pooler_res <= 56")
        w3 = sd[hp + "ASPP.convs.3.0.weight"].float()[:, :, 1:2, 1:2]
        out[hp + "ASPP.convs.3.0"] = _pack_conv(w3, None, device); gn(hp + "ASPP.convs.3.1")
        conv(hp + "ASPP.convs.4.1", with_bias=False); gn(hp + "ASPP.convs.4.2")
        conv(hp + "ASPP.project.0", with_bias=False)
        for i in range(8):
            conv(hp + f"body_conv_fcn{i + 1}", with_bias=False); gn(hp + f"body_conv_fcn{i + 1}.norm")
    else:
        for i in range(8):
            conv(hp + f"body_conv_fcn{i + 1}")
    pp = "roi_heads.densepose_predictor."
    # coarse / fine / u / v, then the confidence heads a WC* model carries (chart_with_confidence.py:50-89); all of them
    # ConvTranspose2d(512, C, 4, stride 2, pad 1), fused on Cout
    names = ("ann_index_lowres", "index_uv_lowres", "u_lowres", "v_lowres") + tuple(h + "_lowres" for h, _ in spec.extra_heads)
    wt = torch.cat([sd[pp + n + ".weight"].float() for n in names], dim=1)     # [512, Ctot, 4, 4]
    bt = torch.cat([sd[pp + n + ".bias"].float() for n in names], dim=0)
    for py in range(2):
        for px in range(2):
            kys = [3 - 2 * t for t in range(2)] if py == 0 else [2 - 2 * t for t in range(2)]
            kxs = [3 - 2 * t for t in range(2)] if px == 0 else [2 - 2 * t for t in range(2)]
            sub = wt[:, :, kys, :][:, :, :, kxs]                                 # [ci, co, ty, tx]
            out[pp + f"phase{py * 2 + px}"] = _pack_khwc(sub.permute(1, 2, 3, 0), bt, device, strict)
    # the four phases stacked on Cout: one GEMM launch whose N blocks are the phases (conv_igemm phase_taps)
    ph = [out.pop(pp + f"phase{i}") for i in range(4)]
    out[pp + "phases"] = (torch.cat([q[0] for q in ph], 0).contiguous(), torch.cat([q[1] for q in ph], 0).contiguous(),
                          ph[0][2], 4 * ph[0][3])
    return out
