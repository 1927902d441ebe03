"""Generates tests/golden/c2_names.json: Caffe2 / Detectron1 blob names of a DensePose R-CNN (R50-FPN) and the
detectron2 names the REFERENCE's converter (detectron2/checkpoint/c2_model_loading.py) gives them.

Run in the build container only (the GPU box has no /root/reference):
    python tests/golden/make_c2_names.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "shims"))
sys.path.insert(0, "/root/reference")

import torch  # noqa: E402


def blob_names(blocks=(3, 4, 6, 3)):
    names = ["conv1_w", "res_conv1_bn_s", "res_conv1_bn_b"]
    for si, nb in enumerate(blocks):
        for bi in range(nb):
            branches = ["branch2a", "branch2b", "branch2c"] + (["branch1"] if bi == 0 else [])
            for br in branches:
                p = f"res{si + 2}_{bi}_{br}"
                names += [p + "_w", p + "_bn_s", p + "_bn_b"]
    last = {2: blocks[0] - 1, 3: blocks[1] - 1, 4: blocks[2] - 1, 5: blocks[3] - 1}
    for lvl in (5, 4, 3, 2):
        lat = f"fpn_inner_res{lvl}_{last[lvl]}_sum" + ("" if lvl == 5 else "_lateral")
        out = f"fpn_res{lvl}_{last[lvl]}_sum"
        names += [lat + "_w", lat + "_b", out + "_w", out + "_b"]
    for p in ("conv_rpn_fpn2", "rpn_cls_logits_fpn2", "rpn_bbox_pred_fpn2", "fc6", "fc7", "cls_score", "bbox_pred"):
        names += [p + "_w", p + "_b"]
    for i in range(1, 9):
        names += [f"body_conv_fcn{i}_w", f"body_conv_fcn{i}_b"]
    for p in ("AnnIndex_lowres", "Index_UV_lowres", "U_lowres", "V_lowres"):
        names += [p + "_w", p + "_b"]
    return names


def main():
    from detectron2.checkpoint.c2_model_loading import convert_c2_detectron_names
    names = blob_names()
    weights = {n: torch.zeros(8, 2) for n in names}
    new, origin = convert_c2_detectron_names(weights)
    mapping = {origin[k]: k for k in new}
    assert len(mapping) == len(names)
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "c2_names.json")
    with open(out, "w") as f:
        json.dump({"reference": "detectron2/checkpoint/c2_model_loading.py::convert_c2_detectron_names", "map": mapping}, f, indent=0, sort_keys=True)
    print(f"{len(mapping)} names -> {out}")


if __name__ == "__main__":
    main()
