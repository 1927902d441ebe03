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
    uv_confidence: str = ""         # ROI_DENSEPOSE_HEAD.UV_CONFIDENCE: "" (off) | "iid_iso" (WC1: sigma_2 head) |
                                    # "indep_aniso" (WC2: sigma_2, kappa_u, kappa_v heads)  (chart_with_confidence.py:50-79)
    segm_confidence: bool = False   # ROI_DENSEPOSE_HEAD.SEGM_CONFIDENCE.ENABLED (WC*M: fine / coarse segm confidence heads)

    @property
    def blocks(self) -> Tuple[int, int, int, int]:
        return {50: (3, 4, 6, 3), 101: (3, 4, 23, 3)}[self.depth]

    @property
    def out_size(self) -> int:
        return 4 * self.pooler_res

    @property
    def extra_heads(self) -> Tuple[Tuple[str, int], ...]:
        """Confidence heads the predictor carries beside coarse / fine / u / v, in the engine's channel order
        (chart_with_confidence.py:50-89): (name, channels)."""
        heads = []
        if self.uv_confidence:
            heads.append(("sigma_2", 25))
            if self.uv_confidence == "indep_aniso":
                heads += [("kappa_u", 25), ("kappa_v", 25)]
        if self.segm_confidence:
            heads += [("fine_segm_confidence", 1), ("coarse_segm_confidence", 1)]
        return tuple(heads)


EXTRA_HEAD_ORDER = ("sigma_2", "kappa_u", "kappa_v", "fine_segm_confidence", "coarse_segm_confidence")


def _mk(name, depth, head, decoder, res, coarse, uv="", segm=False) -> ModelSpec:
    return ModelSpec(name=name, depth=depth, head=head, decoder_on=decoder, pooler_res=res, coarse_ch=coarse,
                     uv_confidence=uv, segm_confidence=segm)


BUILTIN: Dict[str, ModelSpec] = {s.name: s for s in [
    _mk("densepose_rcnn_R_50_FPN_s1x_legacy", 50, "v1convx", False, 14, 15),
    _mk("densepose_rcnn_R_101_FPN_s1x_legacy", 101, "v1convx", False, 14, 15),
    _mk("densepose_rcnn_R_50_FPN_s1x", 50, "v1convx", True, 28, 2),
    _mk("densepose_rcnn_R_101_FPN_s1x", 101, "v1convx", True, 28, 2),
    _mk("densepose_rcnn_R_50_FPN_DL_s1x", 50, "deeplab", True, 28, 2),
    _mk("densepose_rcnn_R_101_FPN_DL_s1x", 101, "deeplab", True, 28, 2),
    # two of the confidence ("WC") variants (README.md:208-330 of the reference): same path, extra predictor heads
    _mk("densepose_rcnn_R_50_FPN_WC1_s1x", 50, "v1convx", True, 28, 2, "iid_iso", False),
    _mk("densepose_rcnn_R_50_FPN_WC2M_s1x", 50, "v1convx", True, 28, 2, "indep_aniso", True),
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


def _require(cond: bool, what: str):
    if not cond:
        raise ValueError("unsupported config: " + what)


def spec_from_yaml(path: str, min_score: float = 0.3, nms_thresh: float = None) -> ModelSpec:
    """export.py:21-33 semantics: yaml (+_BASE_) over the code defaults, then --min_score / --nms_thresh.
    Every key that changes the numerics of the hot path is read and either honoured or rejected: a yaml this engine
    does not implement (other backbone / pooler / predictor / head) raises instead of exporting a model that computes
    something else."""
    if not os.path.isfile(path):
        raise FileNotFoundError(f"config file {path!r} does not exist (builtin names: {', '.join(sorted(BUILTIN))})")
    y = _load_yaml_with_base(path)
    model = y.get("MODEL", {})
    dp = model.get("ROI_DENSEPOSE_HEAD", {})
    name = os.path.splitext(os.path.basename(path))[0]
    head_name = dp.get("NAME", "DensePoseV1ConvXHead")
    heads = {"DensePoseV1ConvXHead": "v1convx", "DensePoseDeepLabHead": "deeplab"}
    _require(head_name in heads, f"MODEL.ROI_DENSEPOSE_HEAD.NAME {head_name}")
    _require(model.get("ROI_HEADS", {}).get("NAME", "DensePoseROIHeads") == "DensePoseROIHeads", "only DensePoseROIHeads models")
    _require(model.get("META_ARCHITECTURE", "GeneralizedRCNN") == "GeneralizedRCNN", "MODEL.META_ARCHITECTURE")
    _require(model.get("BACKBONE", {}).get("NAME", "build_resnet_fpn_backbone") == "build_resnet_fpn_backbone",
             f"MODEL.BACKBONE.NAME {model.get('BACKBONE', {}).get('NAME')}")
    # detectron2/config.py defaults: ROIAlignV2 / sampling 0; the DensePose base yaml sets ROIAlign / 2 for both heads,
    # which is what the kernels implement (aligned=False, two samples per bin axis)
    box_head = model.get("ROI_BOX_HEAD", {})
    _require(box_head.get("POOLER_TYPE", "ROIAlignV2") == "ROIAlign", f"ROI_BOX_HEAD.POOLER_TYPE {box_head.get('POOLER_TYPE', 'ROIAlignV2')}")
    _require(dp.get("POOLER_TYPE", "ROIAlignV2") == "ROIAlign", f"ROI_DENSEPOSE_HEAD.POOLER_TYPE {dp.get('POOLER_TYPE', 'ROIAlignV2')}")
    _require(int(box_head.get("POOLER_SAMPLING_RATIO", 0)) == 2, "ROI_BOX_HEAD.POOLER_SAMPLING_RATIO != 2")
    _require(int(dp.get("POOLER_SAMPLING_RATIO", 2)) == 2, "ROI_DENSEPOSE_HEAD.POOLER_SAMPLING_RATIO != 2")    # densepose/config.py:178
    _require(int(box_head.get("POOLER_RESOLUTION", 14)) == 7, "ROI_BOX_HEAD.POOLER_RESOLUTION != 7")
    _require(box_head.get("NAME", "") == "FastRCNNConvFCHead" and int(box_head.get("NUM_FC", 0)) == 2 and
             int(box_head.get("NUM_CONV", 0)) == 0, "ROI_BOX_HEAD must be FastRCNNConvFCHead with 2 FCs")
    pred = dp.get("PREDICTOR_NAME", "DensePoseChartPredictor")
    _require(pred in ("DensePoseChartPredictor", "DensePoseChartWithConfidencePredictor"), f"PREDICTOR_NAME {pred}")
    _require(int(model.get("ROI_HEADS", {}).get("NUM_CLASSES", 80)) == 1, "ROI_HEADS.NUM_CLASSES != 1")
    _require(not model.get("MASK_ON", False) and not model.get("KEYPOINT_ON", False), "MASK_ON / KEYPOINT_ON")
    _require(bool(model.get("DENSEPOSE_ON", False)), "MODEL.DENSEPOSE_ON is false")
    _require(int(dp.get("NUM_STACKED_CONVS", 8)) == 8 and int(dp.get("CONV_HEAD_DIM", 512)) == 512 and
             int(dp.get("CONV_HEAD_KERNEL", 3)) == 3, "ROI_DENSEPOSE_HEAD stacked convs must be 8 x 3x3 x 512")
    _require(int(dp.get("NUM_PATCHES", 24)) == 24, "ROI_DENSEPOSE_HEAD.NUM_PATCHES != 24")
    _require(int(dp.get("DECONV_KERNEL", 4)) == 4 and int(dp.get("UP_SCALE", 2)) == 2, "DECONV_KERNEL != 4 or UP_SCALE != 2")
    _require(dp.get("DECODER_NORM", "") == "" and int(dp.get("DECODER_CONV_DIMS", 256)) == 256 and
             int(dp.get("DECODER_NUM_CLASSES", 256)) == 256 and int(dp.get("DECODER_COMMON_STRIDE", 4)) == 4, "decoder shape")
    _require(dp.get("DEEPLAB", {}).get("NORM", "GN") == "GN" and int(dp.get("DEEPLAB", {}).get("NONLOCAL_ON", 0)) == 0,
             "DEEPLAB.NORM != GN or NONLOCAL_ON")
    mean = tuple(float(v) for v in model.get("PIXEL_MEAN", (103.530, 116.280, 123.675)))
    std = tuple(float(v) for v in model.get("PIXEL_STD", (1.0, 1.0, 1.0)))
    _require(len(mean) == 3 and len(std) == 3, "PIXEL_MEAN / PIXEL_STD must have 3 entries")
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
        pixel_mean=mean, pixel_std=std,
        uv_confidence=(str(dp.get("UV_CONFIDENCE", {}).get("TYPE", "iid_iso"))
                       if dp.get("UV_CONFIDENCE", {}).get("ENABLED", False) else ""),
        segm_confidence=bool(dp.get("SEGM_CONFIDENCE", {}).get("ENABLED", False)),
    )
    if nms_thresh is not None:
        spec = replace(spec, nms_test=float(nms_thresh))
    _require(spec.depth in (50, 101), f"ResNet depth {spec.depth}")
    _require(spec.input_format in ("BGR", "RGB"), f"INPUT.FORMAT {spec.input_format}")
    _require(spec.pooler_res in (14, 28), f"ROI_DENSEPOSE_HEAD.POOLER_RESOLUTION {spec.pooler_res}")
    _require(spec.uv_confidence in ("", "iid_iso", "indep_aniso"), f"UV_CONFIDENCE.TYPE {spec.uv_confidence}")
    return spec
