"""Weight files -> the reference's state_dict naming (row f2 of SURVEY.md §8).

Covers what `DetectionCheckpointer` does for this path without fvcore / iopath
(detectron2/checkpoint/detection_checkpoint.py:49-122, c2_model_loading.py:10-329):

  * `.pth` / `.pt`  : `torch.load`, unwrap {"model": ...};
  * `.pkl`, detectron2 model-zoo format ({"model": ndarray dict, "__author__": ...}): the six checkpoints of the
    reference's README;
  * `.pkl`, Caffe2 / Detectron1 format ({"blobs": ...} or a flat dict): blob names are rewritten to detectron2 module
    names (`rename_caffe2`), the background row of `cls_score` / `bbox_pred` is moved / dropped, and the result is
    attached to the model's keys by the longest-dotted-suffix rule (`attach_by_suffix`).

The rename is an ordered table of (pattern, replacement) rewrites applied to every key. It is restricted to what a
DensePose R-CNN holds (ResNet + FPN backbone, RPN, box head, DensePose head / predictor); mask and keypoint heads are
outside this engine. tests/test_host.py pins it against the reference's own converter (live when
/root/reference is present, and through tests/golden/c2_names.json everywhere).
"""
import pickle
import re
from typing import Dict, Iterable, List, Mapping, Tuple

import numpy as np
import torch

# (regex, replacement), applied in order to the blob name after '_' -> '.'
_SUFFIX_RULES: List[Tuple[str, str]] = [
    (r"\.b$", ".bias"),
    (r"\.w$", ".weight"),
    # affine-channel / batch-norm / group-norm blobs all become the module's `norm`
    (r"bn\.s$", "norm.weight"), (r"bn\.bias$", "norm.bias"),
    (r"bn\.rm", "norm.running_mean"), (r"bn\.running\.mean$", "norm.running_mean"),
    (r"bn\.riv$", "norm.running_var"), (r"bn\.running\.var$", "norm.running_var"),
    (r"bn\.gamma$", "norm.weight"), (r"bn\.beta$", "norm.bias"),
    (r"gn\.s$", "norm.weight"), (r"gn\.bias$", "norm.bias"),
]
_BACKBONE_RULES: List[Tuple[str, str]] = [
    (r"^res\.conv1\.norm\.", "conv1.norm."),      # the stem's affine blob is named after "res"
    (r"^conv1\.", "stem.conv1."),
    (r"\.branch1\.", ".shortcut."), (r"\.branch2a\.", ".conv1."), (r"\.branch2b\.", ".conv2."), (r"\.branch2c\.", ".conv3."),
]
_DENSEPOSE_RULES: List[Tuple[str, str]] = [
    (r"^body\.conv\.fcn", "body_conv_fcn"),
    (r"AnnIndex\.lowres", "ann_index_lowres"), (r"Index\.UV\.lowres", "index_uv_lowres"),
    (r"U\.lowres", "u_lowres"), (r"V\.lowres", "v_lowres"),
]
_DETECTOR_RULES: List[Tuple[str, str]] = [
    # the FPN RPN head is defined on level 2 and shared, hence "fpn2"; the plain names are the non-FPN models
    (r"conv\.rpn\.fpn2", "proposal_generator.rpn_head.conv"), (r"conv\.rpn", "proposal_generator.rpn_head.conv"),
    (r"rpn\.bbox\.pred\.fpn2", "proposal_generator.rpn_head.anchor_deltas"),
    (r"rpn\.cls\.logits\.fpn2", "proposal_generator.rpn_head.objectness_logits"),
    (r"rpn\.bbox\.pred", "proposal_generator.rpn_head.anchor_deltas"),
    (r"rpn\.cls\.logits", "proposal_generator.rpn_head.objectness_logits"),
    (r"^bbox\.pred", "bbox_pred"), (r"^cls\.score", "cls_score"),
    (r"^fc6\.", "box_head.fc1."), (r"^fc7\.", "box_head.fc2."), (r"^head\.conv", "box_head.conv"),
]
_HARD_CODED = {"pred_b": "linear_b", "pred_w": "linear_w"}


def _fpn_name(name: str) -> str:
    """fpn.inner.res<k>.<n>.sum[.lateral].<p> -> fpn_lateral<k>.<p>;  fpn.res<k>.<n>.sum.<p> -> fpn_output<k>.<p>."""
    parts = name.split(".")
    norm = ".norm" if "norm" in parts else ""
    if name.startswith("fpn.inner."):
        return f"fpn_lateral{int(parts[2][3:])}{norm}.{parts[-1]}"
    if name.startswith("fpn.res"):
        return f"fpn_output{int(parts[1][3:])}{norm}.{parts[-1]}"
    return name


def rename_caffe2_key(key: str) -> str:
    k = _HARD_CODED.get(key, key).replace("_", ".")
    for rules in (_SUFFIX_RULES, _BACKBONE_RULES, _DENSEPOSE_RULES, _DETECTOR_RULES):
        for pat, rep in rules:
            k = re.sub(pat, rep, k)
    return _fpn_name(k)


def rename_caffe2(weights: Mapping[str, torch.Tensor]) -> Tuple[Dict[str, torch.Tensor], Dict[str, str]]:
    """Caffe2 blob dict -> (detectron2-named dict, new name -> blob name). `bbox_pred` loses the four background rows,
    `cls_score` gets its background row moved from index 0 to the end (c2_model_loading.py:182-204)."""
    out: Dict[str, torch.Tensor] = {}
    origin: Dict[str, str] = {}
    for blob in sorted(weights):
        new = rename_caffe2_key(blob)
        if new in out:
            raise ValueError(f"Caffe2 blobs '{origin[new]}' and '{blob}' both map to '{new}'")
        v = torch.as_tensor(weights[blob])
        if new.startswith("bbox_pred."):
            v = v[4:]
        elif new.startswith("cls_score."):
            v = torch.cat([v[1:], v[:1]])
        out[new], origin[new] = v, blob
    return out, origin


def attach_by_suffix(model_shapes: Mapping[str, Iterable[int]], ckpt: Mapping[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """align_and_update_state_dicts (c2_model_loading.py:209-329): a checkpoint key belongs to the model key of which
    it is a whole dotted suffix; the longest such key wins; a shape mismatch skips the pair; one checkpoint key claimed
    by two model keys is an error. Unmatched checkpoint keys are passed through unchanged."""
    ckeys = sorted(ckpt)
    result: Dict[str, torch.Tensor] = {}
    taken: Dict[str, str] = {}
    for mkey in sorted(model_shapes):
        best = ""
        for ck in ckeys:
            if (mkey == ck or mkey.endswith("." + ck)) and len(ck) > len(best):
                best = ck
        if not best:
            continue
        v = ckpt[best]
        if tuple(v.shape) != tuple(model_shapes[mkey]):
            continue
        if best in taken:
            raise ValueError(f"checkpoint key '{best}' matches both '{taken[best]}' and '{mkey}'")
        taken[best] = mkey
        result[mkey] = v
    for ck in ckeys:
        if ck not in taken:
            result[ck] = ckpt[ck]
    return result


def _tensors(d: Mapping[str, object]) -> Dict[str, torch.Tensor]:
    return {k: torch.as_tensor(np.asarray(v)) if not isinstance(v, torch.Tensor) else v for k, v in d.items()}


def load_checkpoint(path: str, model_shapes: Mapping[str, Iterable[int]] = None) -> Dict[str, torch.Tensor]:
    """File -> reference-named state_dict. `model_shapes` (key -> shape of the target model, e.g. from
    `synth.make_state_dict(spec, 0)`) is needed for Caffe2 files, whose names only match by suffix."""
    if path.endswith(".pkl"):
        with open(path, "rb") as f:
            data = pickle.load(f, encoding="latin1")
        if "model" in data and "__author__" in data:
            return _tensors(data["model"])
        blobs = data["blobs"] if "blobs" in data else data
        blobs = {k: v for k, v in blobs.items() if not k.endswith("_momentum")}
        renamed, _ = rename_caffe2(_tensors(blobs))
        if model_shapes is None:
            raise ValueError("a Caffe2-format checkpoint needs the target model's key/shape table to attach its weights")
        return attach_by_suffix(model_shapes, renamed)
    data = torch.load(path, map_location="cpu", weights_only=False)
    if isinstance(data, dict) and isinstance(data.get("model"), dict):
        data = data["model"]
    return data
