"""Runs the bench workload (configs[1]) for a few steps with plain stream launches (no CUDA graph), so that ncu
sees every kernel of a step in launch order. Usage under gpurun:

  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/launches.csv python tools/prof_step.py
  ncu --profile-from-start off --set full --clock-control none --import-source on -o gpurun_out/step_full \
      python tools/prof_step.py

Only the last step sits between cudaProfilerStart/Stop; the `--steps` before it are warm-up.
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

from densepose_torchscript_b200 import synth
from densepose_torchscript_b200.config import BUILTIN
from densepose_torchscript_b200.engine import Engine


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="densepose_rcnn_R_50_FPN_s1x")
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--height", type=int, default=800)
    ap.add_argument("--width", type=int, default=1333)
    ap.add_argument("--steps", type=int, default=2)
    a = ap.parse_args()
    spec = BUILTIN[a.config]
    eng = Engine(spec, synth.make_state_dict(spec, 0), use_graph=False)
    imgs = torch.stack([synth.synthetic_image(a.height, a.width, seed=100 + i) for i in range(a.batch)]).cuda()
    sess = eng.session(a.batch, a.height, a.width, False)
    sys.stderr.write(f"launches per step: {sess.launches}\n")
    for i, (n, _) in enumerate(sess.op_info()):
        sys.stderr.write(f"op {i} {n}\n")
    for _ in range(a.steps):                 # warm-up, not profiled (ncu --profile-from-start off)
        sess.run(imgs)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    sess.run(imgs)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    sys.stderr.write(f"detections: {sess.det_count.cpu().tolist()}\n")


if __name__ == "__main__":
    main()
