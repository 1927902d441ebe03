#!/usr/bin/env python
"""Drop-in for the reference's run.py (run.py:11-64): exported model + image/video -> visualisation.

    python run.py <model.pt> <input image or video> [--cpu] [--fp32]

The exported model's forward pass is libdpb200.so (sm_100a only): `--cpu` is accepted for interface
compatibility but fails loudly — there is no CPU path in this engine.
"""
import argparse
import os

import cv2
import numpy as np
import torch

import densepose_torchscript_b200.torch_ops  # noqa: F401  registers torch.ops.dpb200 before torch.jit.load
from densepose_torchscript_b200.extractor import End2EndVisualizer


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("model", type=str, help="Path to the exported model")
    parser.add_argument("input", type=str, help="Path to the input image or video")
    parser.add_argument("--cpu", action="store_true", help="(unsupported) the engine has no CPU path")
    parser.add_argument("--fp32", action="store_true", help="Emit fp32 outputs on the GPU")
    parser.add_argument("--batch", type=int, default=8, help="video frames per engine batch (the reference is 1)")
    args = parser.parse_args()

    predictor = torch.jit.load(args.model).eval()
    if args.cpu or not torch.cuda.is_available():
        raise SystemExit("dpb200 runs on a B200 only: no CUDA device / --cpu requested")
    predictor = predictor.cuda()
    predictor = predictor.float() if args.fp32 else predictor.half()      # run.py:20-29

    visualizer = End2EndVisualizer(alpha=.7, keep_bg=False)         # run.py:17
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
    n = run_video(cap, predictor, visualizer, save_path, fps, args.batch)
    cap.release()
    print(f"Video ({n} frames) saved to {save_path}")


def run_video(cap, predictor, visualizer, save_path, fps, batch):
    """The reference steps a video frame by frame (run.py:42-57). Same output here, but frames go through the
    engine in batches: uint8 frames -> pinned staging -> device resize/normalise -> forward -> on-device result
    extraction, double-buffered (HostPipeline), so decode / draw on the host overlap the GPU."""
    from densepose_torchscript_b200.engine import HostPipeline
    from densepose_torchscript_b200.torch_ops import engine_of

    eng = engine_of(predictor)
    writer, pipe, n = None, None, 0
    pending = []          # frame batches in flight, oldest first

    def emit(frames, results):
        nonlocal writer, n
        for frame, res in zip(frames, results):
            out = visualizer.draw(frame, res["densepose"], res["boxes_xywh"])
            if writer is None:
                writer = cv2.VideoWriter(save_path, cv2.VideoWriter_fourcc(*"mp4v"), fps, (out.shape[1], out.shape[0]))
            writer.write(out)
            n += 1

    def flush_tail(frames):
        # fewer frames than a batch: one at a time through the exported module, like the reference
        for frame in frames:
            emit([frame], [_single(predictor, visualizer, frame)])

    buf = []
    while True:
        ok, frame = cap.read()
        if ok:
            buf.append(frame)
        if len(buf) == batch or (not ok and buf):
            if len(buf) < batch:
                if pipe is not None:
                    for frames, res in zip(pending, pipe.drain()):
                        emit(frames, res)
                    pending.clear()
                flush_tail(buf)
                buf = []
            else:
                if pipe is None:
                    h, w = buf[0].shape[:2]
                    pipe = HostPipeline(eng, batch, h, w, src_u8=True, depth=2, extract=True, labels_u8=True)
                done = pipe.submit(torch.from_numpy(np.stack(buf)))
                pending.append(buf)
                buf = []
                if done is not None:
                    emit(pending.pop(0), done)
        if not ok:
            break
    if pipe is not None:
        for frames, res in zip(pending, pipe.drain()):
            emit(frames, res)
    if writer is not None:
        writer.release()
    return n


def _single(predictor, visualizer, frame):
    outputs = predictor(torch.from_numpy(frame))
    results, boxes_xywh = visualizer.extractor({k: v.float() if v.is_floating_point() else v for k, v in outputs.items()})
    return {"densepose": results, "boxes_xywh": boxes_xywh}


if __name__ == "__main__":
    main()
