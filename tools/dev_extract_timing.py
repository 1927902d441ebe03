"""Where the batch-1 `extracted` latency goes: forward of a uint8 session, the dp_resample kernel, the D2H of labels / uv.
    python tools/dev_extract_timing.py
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from densepose_torchscript_b200 import ops, synth
from densepose_torchscript_b200.config import BUILTIN
from densepose_torchscript_b200.engine import Engine


def ev_time(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    spec = BUILTIN["densepose_rcnn_R_50_FPN_s1x"]
    eng = Engine(spec, synth.make_state_dict(spec, 0))
    img = synth.synthetic_image(800, 1333, seed=100)
    u8 = img.round().clamp(0, 255).to(torch.uint8)[None].cuda().contiguous()
    f32 = img[None].cuda().contiguous()
    s8, sf = eng.session(1, 800, 1333, True), eng.session(1, 800, 1333, False)
    print("forward u8 session  ms", round(ev_time(lambda: s8.run(u8)), 3))
    print("forward f32 session ms", round(ev_time(lambda: sf.run(f32)), 3))
    res = eng.forward_batch(u8)[0]
    d = len(res["scores"])
    boxes_xywh, wh, offsets = ops.box_sizes(res["pred_boxes"].cpu())
    total = int(offsets[-1])
    print("detections", d, "pixels", total)
    labels = torch.empty(total, dtype=torch.uint8, device="cuda")
    uv = torch.empty(2 * total, dtype=torch.float32, device="cuda")
    wh_d, off_d = wh.cuda(), offsets.cuda()
    c, f, u, v = (res[k].contiguous() for k in ("pred_densepose_coarse_segm", "pred_densepose_fine_segm", "pred_densepose_u", "pred_densepose_v"))
    print("dp_resample kernel  ms", round(ev_time(lambda: ops.dp_resample_into(c, f, u, v, wh_d, off_d, total, labels, uv)), 3))
    hl, hu = torch.empty(total, dtype=torch.uint8).pin_memory(), torch.empty(2 * total, dtype=torch.float32).pin_memory()
    print("D2H labels+uv       ms", round(ev_time(lambda: (hl.copy_(labels, non_blocking=True), hu.copy_(uv, non_blocking=True))), 3),
          "bytes", total * 9)
    for name, t in (("preprocess u8", lambda: ops.preprocess(u8, 800 / 800, spec.pixel_mean, spec.pixel_std)),
                    ("preprocess f32", lambda: ops.preprocess(f32, 800 / 800, spec.pixel_mean, spec.pixel_std))):
        print(name, "ms", round(ev_time(t), 3))
    t0 = time.perf_counter()
    for _ in range(20):
        ops.box_sizes(res["pred_boxes"].cpu())
    print("host box_sizes      ms", round((time.perf_counter() - t0) / 20 * 1e3, 3))


if __name__ == "__main__":
    main()
