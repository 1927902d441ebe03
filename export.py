#!/usr/bin/env python
"""Drop-in for the reference's export.py (export.py:12-41): config + weights -> a TorchScript file.

    python export.py <cfg.yaml | builtin config name> <weights | ""> [--min_score 0.3] [--nms_thresh T] [--fp16]

writes exported/<cfgstem>_fp{32,16}.pt holding the scripted DensePoseB200Predictor (the packed weights and
one custom op, torch.ops.dpb200.forward).  Weights: a torch state_dict (.pth / .pt, optionally under a
"model" key) or a detectron2-format .pkl ({"model": {name: ndarray}, "__author__": ...}); an empty string
uses the seeded synthetic weights (no network here to fetch the model zoo).
"""
import argparse
import os

import torch

from densepose_torchscript_b200 import synth
from densepose_torchscript_b200.config import BUILTIN, spec_from_yaml
from densepose_torchscript_b200.predictor import DensePoseB200Predictor


def load_weights(path: str, spec=None):
    """detectron2/checkpoint/detection_checkpoint.py:49-122 without fvcore / iopath: torch files, detectron2 model-zoo
    pickles and Caffe2 / Detectron1 pickles (blob names converted and attached by suffix, checkpoint.py)."""
    from densepose_torchscript_b200.checkpoint import load_checkpoint
    shapes = None
    if spec is not None and path.endswith(".pkl"):
        shapes = {k: tuple(v.shape) for k, v in synth.make_state_dict(spec, 0).items()}
    return load_checkpoint(path, shapes)


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("cfg", type=str, help="Path to the config file (or a builtin config name)")
    parser.add_argument("weights", type=str, help="Path to the weights file ('' = seeded synthetic weights)")
    parser.add_argument("--min_score", default=0.3, type=float, help="Minimum score threshold (default: %(default)s)")
    parser.add_argument("--nms_thresh", metavar="<threshold>", default=None, type=float, help="NMS threshold")
    parser.add_argument("--fp16", action="store_true", help="Emit fp16 scores / DensePose tensors (like .half())")
    args = parser.parse_args()

    if os.path.exists(args.cfg):
        spec = spec_from_yaml(args.cfg, min_score=args.min_score, nms_thresh=args.nms_thresh)
    else:
        from dataclasses import replace
        if args.cfg not in BUILTIN:
            raise SystemExit(f"{args.cfg!r} is neither a config file nor a builtin config name "
                             f"(builtin: {', '.join(sorted(BUILTIN))})")
        spec = replace(BUILTIN[args.cfg], score_thresh=args.min_score)
        if args.nms_thresh is not None:
            spec = replace(spec, nms_test=args.nms_thresh)
    sd = load_weights(args.weights, spec) if args.weights else synth.make_state_dict(spec, 0)
    predictor = DensePoseB200Predictor(spec, sd).eval()
    predictor = torch.jit.script(predictor)
    if args.fp16:
        predictor = predictor.half()
    os.makedirs("./exported", exist_ok=True)
    stem = os.path.splitext(os.path.basename(args.cfg))[0]
    save_path = os.path.join("./exported", stem + ("_fp16.pt" if args.fp16 else "_fp32.pt"))
    torch.jit.save(predictor, save_path)
    print(f"Model saved to {save_path}")


if __name__ == "__main__":
    main()
