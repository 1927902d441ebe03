"""Model configuration for the hot path.

The reference drives everything from a yacs CfgNode (detectron2/config.py, densepose/config.py,
configs/*.yaml).  Inference needs ~20 of those keys; `ModelSpec` holds them, `spec_from_yaml` reads them
from the reference's own yaml files (following `_BASE_`), and `BUILTIN` lists the six published models
(README.md:69-206) so no yaml is needed for them.
"""
import os
from dataclasses import dataclass, replace
from typing import Dict, Tuple

import yaml


@dataclass(frozen=True)
class ModelSpec:
    name: str
    depth: int = 50                 # MODEL.RESNETS.DEPTH
    head: str = "v1convx"           # MODEL.ROI_DENSEPOSE_HEAD.NAME -> v1convx | deeplab
    decoder_on: bool = True         # MODEL.ROI_DENSEPOSE_HEAD.DECODER_ON
    pooler_res: int = 28            # MODEL.ROI_DENSEPOSE_HEAD.POOLER_RESOLUTION
    coarse_ch: int = 2              # MODEL.ROI_DENSEPOSE_HEAD.NUM_COARSE_SEGM_CHANNELS
    score_thresh: float = 0.3       # MODEL.ROI_HEADS.SCORE_THRESH_TEST (export.py --min_score)
    nms_test: float = 0.5           # MODEL.ROI_HEADS.NMS_THRESH_TEST
    dets_per_image: int = 100       # TEST.DETECTIONS_PER_IMAGE
    min_size: int = 800             # INPUT.MIN_SIZE_TEST
    max_size: int = 1333            # INPUT.MAX_SIZE_TEST
    rpn_pre_topk: int = 1000        # MODEL.RPN.PRE_NMS_TOPK_TEST
    rpn_post_topk: int = 1000       # MODEL.RPN.POST_NMS_TOPK_TEST
    rpn_nms: float = 0.7            # MODEL.RPN.NMS_THRESH
    pixel_mean: Tuple[float, float, float] = (103.530, 116.280, 123.675)
    pixel_std: Tuple[float, float, float] = (1.0, 1.0, 1.0)
    input_format: str = "BGR"       # INPUT.FORMAT

    @property
    def blocks(self) -> Tuple[int, int, int, int]:
        return {50: (3, 4, 6, 3), 101: (3, 4, 23, 3)}[self.depth]

    @property
    def out_size(self) -> int:
        return 4 * self.pooler_res


def _mk(name, depth, head, decoder, res, coarse) -> ModelSpec:
    return ModelSpec(name=name, depth=depth, head=head, decoder_on=decoder, pooler_res=res, coarse_ch=coarse)


BUILTIN: Dict[str, ModelSpec] = {s.name: s for s in [
    _mk("densepose_rcnn_R_50_FPN_s1x_legacy", 50, "v1convx", False, 14, 15),
    _mk("densepose_rcnn_R_101_FPN_s1x_legacy", 101, "v1convx", False, 14, 15),
    _mk("densepose_rcnn_R_50_FPN_s1x", 50, "v1convx", True, 28, 2),
    _mk("densepose_rcnn_R_101_FPN_s1x", 101, "v1convx", True, 28, 2),
    _mk("densepose_rcnn_R_50_FPN_DL_s1x", 50, "deeplab", True, 28, 2),
    _mk("densepose_rcnn_R_101_FPN_DL_s1x", 101, "deeplab", True, 28, 2),
]}


def _load_yaml_with_base(path: str) -> dict:
    with open(path) as f:
        cfg = yaml.safe_load(f) or {}
    base = cfg.pop("_BASE_", None)
    if base is None:
        return cfg
    if not os.path.isabs(base):
        base = os.path.join(os.path.dirname(path), base)
    out = _load_yaml_with_base(base)

    def merge(a, b):
        for k, v in a.items():
            if isinstance(v, dict) and isinstance(b.get(k), dict):
                merge(v, b[k])
            else:
                b[k] = v

    merge(cfg, out)
    return out


def spec_from_yaml(path: str, min_score: float = 0.3, nms_thresh: float = None) -> ModelSpec:
    """export.py:21-33 semantics: yaml (+_BASE_) over the code defaults, then --min_score / --nms_thresh."""
    y = _load_yaml_with_base(path)
    model = y.get("MODEL", {})
    dp = model.get("ROI_DENSEPOSE_HEAD", {})
    name = os.path.splitext(os.path.basename(path))[0]
    head_name = dp.get("NAME", "DensePoseV1ConvXHead")
    heads = {"DensePoseV1ConvXHead": "v1convx", "DensePoseDeepLabHead": "deeplab"}
    if head_name not in heads:
        raise ValueError(f"unsupported ROI_DENSEPOSE_HEAD.NAME {head_name}")
    if model.get("ROI_HEADS", {}).get("NAME", "DensePoseROIHeads") != "DensePoseROIHeads":
        raise ValueError("only DensePoseROIHeads models are supported")
    spec = ModelSpec(
        name=name,
        depth=int(model.get("RESNETS", {}).get("DEPTH", 50)),
        head=heads[head_name],
        decoder_on=bool(dp.get("DECODER_ON", True)),
        pooler_res=int(dp.get("POOLER_RESOLUTION", 28)),          # densepose/config.py:177
        coarse_ch=int(dp.get("NUM_COARSE_SEGM_CHANNELS", 2)),
        score_thresh=float(min_score),
        nms_test=float(model.get("ROI_HEADS", {}).get("NMS_THRESH_TEST", 0.5)),
        rpn_pre_topk=int(model.get("RPN", {}).get("PRE_NMS_TOPK_TEST", 1000)),
        rpn_post_topk=int(model.get("RPN", {}).get("POST_NMS_TOPK_TEST", 1000)),
        rpn_nms=float(model.get("RPN", {}).get("NMS_THRESH", 0.7)),
        min_size=int(y.get("INPUT", {}).get("MIN_SIZE_TEST", 800)),
        max_size=int(y.get("INPUT", {}).get("MAX_SIZE_TEST", 1333)),
        input_format=str(y.get("INPUT", {}).get("FORMAT", "BGR")),
        dets_per_image=int(y.get("TEST", {}).get("DETECTIONS_PER_IMAGE", 100)),
    )
    if nms_thresh is not None:
        spec = replace(spec, nms_test=float(nms_thresh))
    if spec.depth not in (50, 101):
        raise ValueError(f"unsupported ResNet depth {spec.depth}")
    return spec
