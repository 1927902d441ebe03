"""Throughput of K steps on one session vs alternating over 2 / 3 sessions (own streams, graph replay each)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from densepose_torchscript_b200 import synth
from densepose_torchscript_b200.config import BUILTIN
from densepose_torchscript_b200.engine import Engine

spec = BUILTIN["densepose_rcnn_R_50_FPN_s1x"]
eng = Engine(spec, synth.make_state_dict(spec, 0))
B, H, W = 8, 800, 1333
imgs = torch.stack([synth.synthetic_image(H, W, seed=100 + i) for i in range(B)]).cuda()
for nsess in (1, 2, 3, 1, 2):
    sessions = [eng.session(B, H, W, False, slot=i) for i in range(nsess)]
    for s in sessions:
        for _ in range(3):
            s.run(imgs)
    torch.cuda.synchronize()
    K = 24
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cur = torch.cuda.current_stream()
    e0.record()
    for i in range(K):
        sessions[i % nsess].run(imgs)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"sessions {nsess}: {ms / K:.3f} ms/step  {B * K / ms * 1e3:.1f} images/s")
