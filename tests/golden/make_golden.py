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
# extra fixtures (name -> (config, image, uint8 input)): the uint8 image path of run.py:33-36 (ATen resizes uint8 in
# fixed point) and the two benchmarked shapes at full size (BASELINE configs[1] and configs[3])
EXTRA = {
    "densepose_rcnn_R_50_FPN_s1x__u8": ("densepose_rcnn_R_50_FPN_s1x", dict(height=240, width=600, seed=3), True),
    "densepose_rcnn_R_50_FPN_s1x_legacy__u8": ("densepose_rcnn_R_50_FPN_s1x_legacy", dict(height=240, width=600, seed=3), True),
    "densepose_rcnn_R_50_FPN_s1x__800x1333": ("densepose_rcnn_R_50_FPN_s1x", dict(height=800, width=1333, seed=1), False),
    "densepose_rcnn_R_101_FPN_s1x__1080p": ("densepose_rcnn_R_101_FPN_s1x", dict(height=1080, width=1920, seed=11), False),
    "densepose_rcnn_R_101_FPN_s1x__1080p_u8": ("densepose_rcnn_R_101_FPN_s1x", dict(height=1080, width=1920, seed=11), True),
    # confidence variants: the reference builds sigma_2 / kappa_u / kappa_v / segm-confidence layers but its forward drops
    # them (chart_with_confidence.py:91-109); the fixture also holds those layers applied to the reference's own head
    # output (interp2d(layer(head_outputs)), as upstream DensePose emits them)
    "densepose_rcnn_R_50_FPN_WC1_s1x": ("densepose_rcnn_R_50_FPN_WC1_s1x", dict(height=240, width=600, seed=3), False),
    "densepose_rcnn_R_50_FPN_WC2M_s1x": ("densepose_rcnn_R_50_FPN_WC2M_s1x", dict(height=240, width=600, seed=3), False),
}
DP_KEYS = ["pred_densepose_coarse_segm", "pred_densepose_fine_segm", "pred_densepose_u", "pred_densepose_v"]
N_LABELS = 16      # detections whose native-resolution part-label map is stored
N_SAMPLE = 8       # detections whose strided DensePose samples are stored (fp16)


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


def make_image(image: dict, u8: bool):
    """The seeded synthetic image; u8: rounded to uint8 HWC, what cv2.imread hands run.py:33-36."""
    img = W.synthetic_image(**image)
    return img.round().clamp(0, 255).to(torch.uint8) if u8 else img


def label_map(out, n):
    """Native-resolution (4S x 4S) part labels of the first n detections: argmax25(fine) * (argmaxK(coarse) > 0)."""
    fine = out["pred_densepose_fine_segm"][:n].argmax(1)
    fg = out["pred_densepose_coarse_segm"][:n].argmax(1) > 0
    return (fine * fg).to(torch.uint8)


def summarize(out):
    fx = {"pred_boxes": out["pred_boxes"].clone(), "scores": out["scores"].clone(),
          "pred_classes": out["pred_classes"].clone(), "image_size": out["image_size"].clone()}
    for k in DP_KEYS:
        t = out[k]
        fx[k + ".shape"] = tuple(t.shape)
        fx[k + ".sum"] = float(t.double().sum())
        fx[k + ".abssum"] = float(t.double().abs().sum())
        fx[k + ".sample"] = t[:4, :, ::8, ::8].clone()
        fx[k + ".sample16"] = t[:N_SAMPLE, :, ::8, ::8].half()
    fx["labels_native"] = label_map(out, N_LABELS)
    return fx


def generate(name: str, config: str, image: dict, u8: bool, here: str):
    spec = O.SPECS[config]
    sd = W.make_state_dict(spec, 0)
    pred = build_reference(config)
    pred.load_state_dict(W.add_aliases(sd, spec), strict=True)
    img = make_image(image, u8)
    captured = {}
    predictor = pred.model.roi_heads.densepose_predictor
    hook = predictor.register_forward_hook(lambda mod, inp, outp: captured.update(head=inp[0]))
    with torch.no_grad():
        out = pred(img)
        hook.remove()
        for head, _ in spec.extra_heads:      # every detection survives detector_postprocess here (asserted below)
            out["pred_densepose_" + head] = predictor.interp2d(getattr(predictor, head + "_lowres")(captured["head"]))
            assert len(out["pred_densepose_" + head]) == len(out["scores"])
    fx = summarize(out)
    for head, _ in spec.extra_heads:
        fx["pred_densepose_" + head + ".sample16"] = out["pred_densepose_" + head][:N_SAMPLE, :, ::8, ::8].half()
    fx["extra_heads"] = [h for h, _ in spec.extra_heads]
    fx["config"] = config
    fx["image"] = dict(image)
    fx["uint8"] = u8
    fx["torch"] = torch.__version__
    fx["threads"] = torch.get_num_threads()        # > 1: ATen's separable float resize kernel (oracle/aten_interp.py)
    # the visualizer's per-box extractor on the first three detections (visualizer.py:46-56)
    from visualizer import DensePoseResultExtractor
    sub = {k: (v[:3] if k != "image_size" else v) for k, v in out.items()}
    results, _ = DensePoseResultExtractor()(sub)
    fx["extract.labels"] = [r["labels"].to(torch.uint8) for r in results]
    fx["extract.uv_sum"] = [float(r["uv"].double().sum()) for r in results]
    path = os.path.join(here, name + ".pt")
    torch.save(fx, path)
    print(name, "detections", len(out["scores"]), "->", path, os.path.getsize(path) // 1024, "KiB")


def main():
    here = os.path.dirname(os.path.abspath(__file__))
    assert torch.get_num_threads() > 1, "generate with > 1 intra-op threads (the float resize kernel depends on it)"
    todo = sys.argv[1:] or (CONFIGS + list(EXTRA))
    for name in todo:
        if name in EXTRA:
            config, image, u8 = EXTRA[name]
            generate(name, config, image, u8, here)
        else:
            generate(name, name, IMAGE, False, here)


if __name__ == "__main__":
    main()
