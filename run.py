#!/usr/bin/env python
"""Drop-in for the reference's run.py (run.py:11-64): exported model + image/video -> visualisation.

    python run.py <model.pt> <input image or video> [--cpu] [--fp32]

The exported model's forward pass is libdpb200.so (sm_100a only): `--cpu` is accepted for interface
compatibility but fails loudly — there is no CPU path in this engine.
"""
import argparse
import os

import cv2
import torch

import densepose_torchscript_b200.torch_ops  # noqa: F401  registers torch.ops.dpb200 before torch.jit.load
from densepose_torchscript_b200.extractor import End2EndVisualizer


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("model", type=str, help="Path to the exported model")
    parser.add_argument("input", type=str, help="Path to the input image or video")
    parser.add_argument("--cpu", action="store_true", help="(unsupported) the engine has no CPU path")
    parser.add_argument("--fp32", action="store_true", help="Emit fp32 outputs on the GPU")
    args = parser.parse_args()

    predictor = torch.jit.load(args.model).eval()
    if args.cpu or not torch.cuda.is_available():
        raise SystemExit("dpb200 runs on a B200 only: no CUDA device / --cpu requested")
    predictor = predictor.cuda()
    predictor = predictor.float() if args.fp32 else predictor.half()      # run.py:20-29

    visualizer = End2EndVisualizer(alpha=.7, inplace=True)
    save_path = "_pred".join(os.path.splitext(args.input))
    img = cv2.imread(args.input)
    if img is not None:
        outputs = predictor(torch.from_numpy(img))                           # uint8 HWC BGR, like run.py:33-36
        cv2.imwrite(save_path, visualizer.visualize(img, outputs))
        print(f"Image saved to {save_path}")
        return
    cap = cv2.VideoCapture(args.input)
    if not cap.isOpened():
        raise SystemExit(f"cannot read {args.input} as an image or a video")
    fps = cap.get(cv2.CAP_PROP_FPS) or 25.0
    writer = None
    n = 0
    while True:
        ok, frame = cap.read()
        if not ok:
            break
        outputs = predictor(torch.from_numpy(frame))
        frame = visualizer.visualize(frame, outputs)
        if writer is None:
            writer = cv2.VideoWriter(save_path, cv2.VideoWriter_fourcc(*"mp4v"), fps, (frame.shape[1], frame.shape[0]))
        writer.write(frame)
        n += 1
    cap.release()
    if writer is not None:
        writer.release()
    print(f"Video ({n} frames) saved to {save_path}")


if __name__ == "__main__":
    main()
