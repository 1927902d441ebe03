"""Bring-up: engine vs oracle (bf16 numerics mode) stage by stage on one image. Run under gpurun."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

from densepose_torchscript_b200.config import BUILTIN
from densepose_torchscript_b200.engine import Engine
from oracle import densepose_oracle as O
from oracle import weights as W


def rel(a: torch.Tensor, b: torch.Tensor):
    a, b = a.float().cpu(), b.float().cpu()
    if a.shape != b.shape:
        return f"SHAPE {tuple(a.shape)} vs {tuple(b.shape)}"
    d = (a - b)
    return "rel_l2=%.3e max_abs=%.3e (ref max %.3e)" % (d.norm() / max(b.norm(), 1e-12), d.abs().max(), b.abs().max())


def nchw(t):  # engine NHWC tap -> NCHW
    return t.permute(0, 3, 1, 2)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="densepose_rcnn_R_50_FPN_s1x")
    ap.add_argument("--h", type=int, default=240)
    ap.add_argument("--w", type=int, default=600)
    ap.add_argument("--batch", type=int, default=1)
    a = ap.parse_args()
    spec_o = O.SPECS[a.config]
    spec = BUILTIN[a.config]
    sd = W.make_state_dict(spec_o, 0)
    img = W.synthetic_image(a.h, a.w, seed=3)
    t0 = time.time()
    taps = {}
    ref = O.forward(img, sd, spec_o, mode="bf16", taps=taps)
    print("oracle bf16 forward %.1fs, detections %d" % (time.time() - t0, len(ref["scores"])))
    eng = Engine(spec, sd)
    imgs = torch.stack([img] * a.batch).cuda()
    sess = eng.session(a.batch, a.h, a.w, False)
    print("session: launches", sess.launches, "workspace MB", sess.workspace.numel() / 1e6, "geometry",
          sess.hr, sess.wr, sess.hp, sess.wp)
    sess.run(imgs)
    torch.cuda.synchronize()
    from densepose_torchscript_b200.ops import stem_to_image
    x0 = stem_to_image(sess.tap("stem_in"))[0]   # [Hp, Wp + 8, 4]
    print("stem_in", rel(x0[:, 4:4 + sess.wp, :3].permute(2, 0, 1), taps["images"][0]))
    for k in ("res2", "res3", "res4", "res5"):
        print(k, rel(nchw(sess.tap(k))[:1], taps["res"][k]))
    for k in ("p2", "p3", "p4", "p5"):
        print(k, rel(nchw(sess.tap(k))[:1], taps["feats"][k]))
    for l in range(5):
        h = sess.tap(f"rpn_head{l}")[0]
        print(f"rpn logits{l}", rel(h[..., :3].reshape(-1), taps["rpn_logits"][l][0]))
        print(f"rpn deltas{l}", rel(h[..., 3:15].reshape(-1, 4), taps["rpn_deltas"][l][0]))
    n = int(sess.tap("proposal_count")[0, 0, 0, 0])
    pb = sess.tap("proposal_boxes")[0, :n, :, 0]
    rp = taps["proposals"]["proposal_boxes"]
    print("proposals n", n, "ref", len(rp))
    m = min(n, len(rp))
    print("proposal boxes (first %d)" % m, rel(pb[:m], rp[:m]))
    same = (pb[:m].cpu() - rp[:m]).abs().max(dim=1).values < 0.5
    print("proposal rows within 0.5px: %d / %d" % (int(same.sum()), m))
    bp = sess.tap("box_pooled")[:n]
    print("box_pooled", rel(bp.permute(0, 3, 1, 2)[:m], taps["box_pooled"][:m]))
    bh = sess.tap("box_head_out")[:n, 0, 0]
    print("box cls", rel(bh[:m, :2], taps["box_cls"][:m]))
    print("box deltas", rel(bh[:m, 2:6], taps["box_deltas"][:m]))
    res = sess.results()[0]
    D, Dr = len(res["scores"]), len(ref["scores"])
    print("detections", D, "ref", Dr)
    md = min(D, Dr)
    print("scores", rel(res["scores"][:md], ref["scores"][:md]))
    print("pred_boxes", rel(res["pred_boxes"][:md], ref["pred_boxes"][:md]))
    if "decoder" in taps:
        print("decoder", rel(nchw(sess.tap("decoder"))[:1], taps["decoder"]))
    print("dp_pooled", rel(nchw(sess.tap("dp_pooled"))[:md], taps["dp_pooled"][:md]))
    print("dp_head", rel(nchw(sess.tap("dp_head"))[:md], taps["dp_head"][:md]))
    for k in ("coarse_segm", "fine_segm", "u", "v"):
        print(k, rel(res["pred_densepose_" + k][:md], ref["pred_densepose_" + k][:md]))
    lab_e = res["pred_densepose_fine_segm"][:md].argmax(1).cpu()
    lab_r = ref["pred_densepose_fine_segm"][:md].argmax(1)
    print("fine label agreement %.5f" % (lab_e == lab_r).float().mean().item())
    # timing
    for _ in range(3):
        sess.run(imgs)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        sess.run(imgs)
    e1.record()
    torch.cuda.synchronize()
    print("engine ms/forward (batch %d): %.3f" % (a.batch, e0.elapsed_time(e1) / 10))


if __name__ == "__main__":
    main()
