"""Generates tests/golden/*.pt by running the REAL reference (/root/reference, through oracle/shims).

Run in the build container only (the GPU box has no /root/reference):
    python tests/golden/make_golden.py
Each fixture holds, for one config, the reference's outputs on a seeded synthetic image with the seeded
calibrated weights of oracle/weights.py: boxes, scores, detection count, and for the four DensePose
tensors float64 checksums plus a strided sample of the first detections (kept small on purpose).
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "shims"))
sys.path.insert(0, "/root/reference")

import torch  # noqa: E402

from oracle import densepose_oracle as O  # noqa: E402
from oracle import weights as W  # noqa: E402

# all six published configs (README.md:69-206 of the reference); `python make_golden.py <name> ...` regenerates a subset
CONFIGS = ["densepose_rcnn_R_50_FPN_s1x_legacy", "densepose_rcnn_R_50_FPN_s1x", "densepose_rcnn_R_101_FPN_DL_s1x",
           "densepose_rcnn_R_50_FPN_DL_s1x", "densepose_rcnn_R_101_FPN_s1x", "densepose_rcnn_R_101_FPN_s1x_legacy"]
IMAGE = dict(height=240, width=600, seed=3)
DP_KEYS = ["pred_densepose_coarse_segm", "pred_densepose_fine_segm", "pred_densepose_u", "pred_densepose_v"]


def build_reference(name: str):
    from detectron2.config import get_cfg
    from detectron2.engine.defaults import DefaultPredictor
    from densepose import add_densepose_config

    cfg = get_cfg()
    add_densepose_config(cfg)
    cfg.merge_from_file(f"/root/reference/configs/{name}.yaml")
    cfg.merge_from_list(["MODEL.ROI_HEADS.SCORE_THRESH_TEST", "0.3"])   # export.py:23-24
    cfg.MODEL.WEIGHTS = ""
    cfg.freeze()
    return DefaultPredictor(cfg)


def summarize(out):
    fx = {"pred_boxes": out["pred_boxes"].clone(), "scores": out["scores"].clone(),
          "pred_classes": out["pred_classes"].clone(), "image_size": out["image_size"].clone()}
    for k in DP_KEYS:
        t = out[k]
        fx[k + ".shape"] = tuple(t.shape)
        fx[k + ".sum"] = float(t.double().sum())
        fx[k + ".abssum"] = float(t.double().abs().sum())
        fx[k + ".sample"] = t[:4, :, ::8, ::8].clone()
    return fx


def main():
    here = os.path.dirname(os.path.abspath(__file__))
    for name in (sys.argv[1:] or CONFIGS):
        spec = O.SPECS[name]
        sd = W.make_state_dict(spec, 0)
        pred = build_reference(name)
        pred.load_state_dict(W.add_aliases(sd, spec), strict=True)
        img = W.synthetic_image(**IMAGE)
        with torch.no_grad():
            out = pred(img)
        fx = summarize(out)
        fx["config"] = name
        fx["image"] = dict(IMAGE)
        fx["torch"] = torch.__version__
        # the visualizer's per-box extractor on the first three detections (visualizer.py:46-56)
        from visualizer import DensePoseResultExtractor
        sub = {k: (v[:3] if k != "image_size" else v) for k, v in out.items()}
        results, _ = DensePoseResultExtractor()(sub)
        fx["extract.labels"] = [r["labels"].to(torch.uint8) for r in results]
        fx["extract.uv_sum"] = [float(r["uv"].double().sum()) for r in results]
        path = os.path.join(here, name + ".pt")
        torch.save(fx, path)
        print(name, "detections", len(out["scores"]), "->", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
