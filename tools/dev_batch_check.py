"""Bring-up: where does a batched run diverge from the single-image run? (run under gpurun)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from densepose_torchscript_b200.config import BUILTIN
from densepose_torchscript_b200.engine import Engine
from oracle import densepose_oracle as O
from oracle import weights as W

name = "densepose_rcnn_R_50_FPN_s1x"
sd = W.make_state_dict(O.SPECS[name], 0)
eng = Engine(BUILTIN[name], sd)
a, b = W.synthetic_image(240, 600, seed=3), W.synthetic_image(240, 600, seed=4)
s1 = eng.session(1, 240, 600, False)
names = ["stem_in", "stem_pool", "res2", "res3", "res4", "res5", "p5", "p4", "p3", "p2", "rpn_head0", "rpn_head1",
         "rpn_head2", "rpn_head3", "rpn_head4", "rpn_cand_scores", "rpn_cand_boxes", "rpn_cand_keep", "proposal_boxes",
         "box_pooled", "box_head_out", "det_boxes_raw", "decoder", "dp_pooled", "dp_head", "dp_lowres"]
singles = []
for img in (a, b):
    s1.run(img[None].cuda().contiguous())
    torch.cuda.synchronize()
    singles.append({n: s1.tap(n).clone() for n in names})
s3 = eng.session(3, 240, 600, False)
s3.run(torch.stack([a, b, a]).cuda().contiguous())
torch.cuda.synchronize()
for n in names:
    t = s3.tap(n)
    for bi, si in ((0, 0), (1, 1), (2, 0)):
        ref = singles[si][n]
        per = t.shape[0] // 3
        got = t[bi * per:(bi + 1) * per]
        if got.shape != ref.shape:
            print(n, bi, "shape", tuple(got.shape), tuple(ref.shape))
            continue
        eq = torch.equal(got, ref)
        if not eq:
            d = (got.float() - ref.float()).abs()
            print(f"{n} image {bi}: DIFF max {float(d.max()):.4g} frac_neq {float((d > 0).float().mean()):.4g}")
        else:
            print(f"{n} image {bi}: equal")
