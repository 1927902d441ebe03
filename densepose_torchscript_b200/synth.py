"""Seeded, well-conditioned random weights and images for benchmarks, smoke tests and parity tests.

The reference's default random init is numerically degenerate (FrozenBN is the identity,
activations reach 1e3, every proposal collapses to a clipped sliver), so synthetic runs use:
  weight = seeded N(0, 2/fan_in) * scale[layer],  FrozenBN / GroupNorm affine + running stats seeded, O(1).
`scale[layer]` is one scalar per conv/linear layer that makes the layer's pre-norm output std hit a target
on a seeded calibration image.  The scalars were computed ONCE by the test-side calibration pass
(python -m oracle.weights --calibrate) and are committed as synth_weight_scales.json, so every machine
regenerates bit-identical tensors with no data-dependent step.

Key names follow the reference modules: backbone/resnet.py:401-403, fpn.py:92-100, rpn.py:108-125,
box_head.py:66-72, fast_rcnn.py:200-203, roi_head.py:29-69, v1convx.py:38-41, deeplab.py:34-60,112-139,
chart.py:45-59.  `add_aliases` adds the duplicate registrations the reference's state_dict carries
 so `load_state_dict(strict=True)` works on the real reference.
"""
import json
import math
import os
from typing import Dict, List, Optional, Tuple

import torch

from .config import ModelSpec

SCALES_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "synth_weight_scales.json")

# pre-norm output std targets (everything else: 1.0)
TARGETS = {
    "proposal_generator.rpn_head.objectness_logits": 2.0,
    "proposal_generator.rpn_head.anchor_deltas": 0.6,
    "roi_heads.box_predictor.cls_score": 1.5,
    "roi_heads.box_predictor.bbox_pred": 0.5,
}


def layer_table(spec: ModelSpec) -> List[Tuple[str, str, Tuple[int, ...]]]:
    """Ordered (prefix, kind, weight shape). kind: conv_bn | conv_bias | conv | conv_gn | linear | deconv | gn."""
    t: List[Tuple[str, str, Tuple[int, ...]]] = []
    bu = "backbone.bottom_up."
    t.append((bu + "stem.conv1", "conv_bn", (64, 3, 7, 7)))
    cin = 64
    for si, nb in enumerate(spec.blocks):
        bott, cout = 64 * 2 ** si, 256 * 2 ** si
        for bi in range(nb):
            p = f"{bu}res{si + 2}.{bi}"
            if bi == 0:
                t.append((p + ".shortcut", "conv_bn", (cout, cin, 1, 1)))
            t.append((p + ".conv1", "conv_bn", (bott, cin, 1, 1)))
            t.append((p + ".conv2", "conv_bn", (bott, bott, 3, 3)))
            t.append((p + ".conv3", "conv_bn", (cout, bott, 1, 1)))
            cin = cout
    for lvl, c in ((2, 256), (3, 512), (4, 1024), (5, 2048)):
        t.append((f"backbone.fpn_lateral{lvl}", "conv_bias", (256, c, 1, 1)))
        t.append((f"backbone.fpn_output{lvl}", "conv_bias", (256, 256, 3, 3)))
    rp = "proposal_generator.rpn_head."
    t.append((rp + "conv", "conv_bias", (256, 256, 3, 3)))
    t.append((rp + "objectness_logits", "conv_bias", (3, 256, 1, 1)))
    t.append((rp + "anchor_deltas", "conv_bias", (12, 256, 1, 1)))
    t.append(("roi_heads.box_head.fc1", "linear", (1024, 256 * 7 * 7)))
    t.append(("roi_heads.box_head.fc2", "linear", (1024, 1024)))
    t.append(("roi_heads.box_predictor.cls_score", "linear", (2, 1024)))
    t.append(("roi_heads.box_predictor.bbox_pred", "linear", (4, 1024)))
    if spec.decoder_on:
        for name, n in (("p2", 1), ("p3", 1), ("p4", 2), ("p5", 3)):
            for k in range(n):
                t.append((f"roi_heads.decoder.{name}.{k if name == 'p2' else 2 * k}", "conv_bias", (256, 256, 3, 3)))
        t.append(("roi_heads.decoder.predictor", "conv_bias", (256, 256, 1, 1)))
    hp = "roi_heads.densepose_head."
    if spec.head == "deeplab":
        t.append((hp + "ASPP.convs.0.0", "conv", (256, 256, 1, 1)))
        t.append((hp + "ASPP.convs.0.1", "gn", (256,)))
        for i in (1, 2, 3):
            t.append((hp + f"ASPP.convs.{i}.0", "conv", (256, 256, 3, 3)))
            t.append((hp + f"ASPP.convs.{i}.1", "gn", (256,)))
        t.append((hp + "ASPP.convs.4.1", "conv", (256, 256, 1, 1)))
        t.append((hp + "ASPP.convs.4.2", "gn", (256,)))
        t.append((hp + "ASPP.project.0", "conv", (256, 1280, 1, 1)))
        for i in range(8):
            t.append((hp + f"body_conv_fcn{i + 1}", "conv_gn", (512, 256 if i == 0 else 512, 3, 3)))
    else:
        for i in range(8):
            t.append((hp + f"body_conv_fcn{i + 1}", "conv_bias", (512, 256 if i == 0 else 512, 3, 3)))
    pp = "roi_heads.densepose_predictor."
    t.append((pp + "ann_index_lowres", "deconv", (512, spec.coarse_ch, 4, 4)))
    t.append((pp + "index_uv_lowres", "deconv", (512, 25, 4, 4)))
    t.append((pp + "u_lowres", "deconv", (512, 25, 4, 4)))
    t.append((pp + "v_lowres", "deconv", (512, 25, 4, 4)))
    for head, ch in getattr(spec, "extra_heads", ()):          # WC* confidence heads (chart_with_confidence.py:50-89)
        t.append((pp + head + "_lowres", "deconv", (512, ch, 4, 4)))
    return t


def load_scales(spec: ModelSpec) -> Optional[Dict[str, float]]:
    if not os.path.exists(SCALES_PATH):
        return None
    import re
    with open(SCALES_PATH) as f:
        allsc = json.load(f)
    # the WC* variants share every calibrated layer with their base config (their extra heads keep scale 1)
    return allsc.get(spec.name) or allsc.get(re.sub(r"_WC\d+M?", "", spec.name))


def make_state_dict(spec: ModelSpec, seed: int = 0, scales: Optional[Dict[str, float]] = None,
                    use_committed_scales: bool = True) -> Dict[str, torch.Tensor]:
    """Canonical (alias-free, prefix-free) state dict."""
    if scales is None and use_committed_scales:
        scales = load_scales(spec)
    scales = scales or {}
    g = torch.Generator().manual_seed(1000003 * seed + 17)
    sd: Dict[str, torch.Tensor] = {}

    def randn(*shape):
        return torch.randn(*shape, generator=g)

    def rand(*shape):
        return torch.rand(*shape, generator=g)

    for prefix, kind, shape in layer_table(spec):
        if kind == "gn":
            sd[prefix + ".weight"] = 0.5 + rand(shape[0])
            sd[prefix + ".bias"] = 0.2 * randn(shape[0])
            continue
        if kind == "deconv":
            fan_in = shape[0] * 4          # each output pixel sees 2x2 taps of every input channel
            cout = shape[1]
        elif kind == "linear":
            fan_in, cout = shape[1], shape[0]
        else:
            fan_in, cout = shape[1] * shape[2] * shape[3], shape[0]
        gain = math.sqrt(2.0 / fan_in)
        target = TARGETS.get(prefix, 1.0)
        sd[prefix + ".weight"] = randn(*shape) * gain * float(scales.get(prefix, 1.0))
        if kind == "conv_bn":
            sd[prefix + ".norm.weight"] = 0.5 + rand(cout)
            sd[prefix + ".norm.bias"] = 0.2 * randn(cout)
            sd[prefix + ".norm.running_mean"] = 0.1 * randn(cout)
            sd[prefix + ".norm.running_var"] = 0.5 + rand(cout)
        elif kind == "conv_gn":
            sd[prefix + ".norm.weight"] = 0.5 + rand(cout)
            sd[prefix + ".norm.bias"] = 0.2 * randn(cout)
        elif kind in ("conv_bias", "linear", "deconv"):
            sd[prefix + ".bias"] = 0.1 * target * randn(cout)
    # spread proposals over FPN levels: grow anchors (dw, dh channels a*4+2, a*4+3)
    b = sd["proposal_generator.rpn_head.anchor_deltas.bias"]
    for a in range(3):
        b[a * 4 + 2] += 1.5
        b[a * 4 + 3] += 1.5
    return sd


def add_aliases(sd: Dict[str, torch.Tensor], spec: ModelSpec, prefix: str = "model.") -> Dict[str, torch.Tensor]:
    """Reference state_dict key set: canonical keys + duplicate registrations, under `prefix`."""
    out: Dict[str, torch.Tensor] = {}
    for k, v in sd.items():
        out[k] = v
        if k.startswith("backbone.bottom_up.res"):
            rest = k[len("backbone.bottom_up.res"):]
            stage, tail = rest.split(".", 1)
            out[f"backbone.bottom_up.stages.{int(stage) - 2}.{tail}"] = v          # resnet.py:401-403
        for lvl in (2, 3, 4, 5):
            for kind in ("lateral", "output"):
                src = f"backbone.fpn_{kind}{lvl}."
                if k.startswith(src):
                    out[f"backbone.{kind}_convs.{5 - lvl}." + k[len(src):]] = v    # fpn.py:92-100 (reversed order)
        if k.startswith("roi_heads.decoder.p"):
            rest = k[len("roi_heads.decoder.p"):]
            if rest[0].isdigit():
                lvl, tail = rest.split(".", 1)
                out[f"roi_heads.decoder.scale_heads.{int(lvl) - 2}.{tail}"] = v     # roi_head.py:66-67
        if k.startswith("roi_heads.densepose_head.body_conv_fcn"):
            rest = k[len("roi_heads.densepose_head.body_conv_fcn"):]
            idx, tail = rest.split(".", 1)
            out[f"roi_heads.densepose_head.stacked_convs.{int(idx) - 1}.{tail}"] = v   # v1convx.py:38-41
    # pixel_mean/std and the cell anchors are non-persistent buffers in the reference (rcnn.py:62-63,
    # anchor_generator.py:25-36): they are not part of its state_dict.
    return {prefix + k: v for k, v in out.items()}


def synthetic_image(height: int = 800, width: int = 1333, seed: int = 1) -> torch.Tensor:
    """Seeded low-frequency image (HWC float32, 0..255, BGR): bicubic-upsampled noise + fine noise."""
    g = torch.Generator().manual_seed(seed)
    low = torch.rand(1, 3, max(height // 32, 2), max(width // 32, 2), generator=g)
    img = torch.nn.functional.interpolate(low, size=(height, width), mode="bicubic", align_corners=False)[0] * 255.0
    img = img + 8.0 * torch.randn(3, height, width, generator=g)
    return img.clamp(0, 255).permute(1, 2, 0).contiguous()


