"""One small forward per numerics mode / input type, for compute-sanitizer:
    compute-sanitizer --tool memcheck python tools/dev_memcheck_forward.py        (also: --tool racecheck)
covers the benchmarked bf16 mode (CTA-pair convs, clustered RPN / NMS kernels), the strict-mode conv path (second A tensor map, hi/lo epilogue), the split stage kernels, the uint8 resize tables,
the WC heads and the DeepLab head in strict mode."""
import os
import sys
from dataclasses import replace

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from densepose_torchscript_b200 import synth
from densepose_torchscript_b200.config import BUILTIN
from densepose_torchscript_b200.engine import Engine

for name, strict, u8 in (("densepose_rcnn_R_50_FPN_s1x", False, False), ("densepose_rcnn_R_50_FPN_s1x", True, False), ("densepose_rcnn_R_50_FPN_DL_s1x", True, True),
                         ("densepose_rcnn_R_50_FPN_WC2M_s1x", False, True), ("densepose_rcnn_R_50_FPN_s1x_legacy", True, False)):
    spec = replace(BUILTIN[name], min_size=192, max_size=320)
    eng = Engine(spec, synth.make_state_dict(spec, 0), strict=strict, use_graph=False)
    img = synth.synthetic_image(96, 160, seed=3)
    if u8:
        img = img.round().clamp(0, 255).to(torch.uint8)
    res = eng.forward_batch(torch.stack([img, img]))
    torch.cuda.synchronize()
    print(name, "strict" if strict else "bf16", "u8" if u8 else "f32", [len(r["scores"]) for r in res], sorted(res[0])[-3:])
