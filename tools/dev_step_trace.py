"""Per-step device time of back-to-back steps (CUDA events around every step), graph replay vs plain launches, and
the sum of per-launch event times of the same session: where does the steady-state step lose time against the sum of
its launches (power-cap clock droop vs launch gaps)?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from densepose_torchscript_b200 import synth
from densepose_torchscript_b200.config import BUILTIN
from densepose_torchscript_b200.engine import Engine

spec = BUILTIN["densepose_rcnn_R_50_FPN_s1x"]
B, H, W = 8, 800, 1333
imgs = torch.stack([synth.synthetic_image(H, W, seed=100 + i) for i in range(B)]).cuda()
for use_graph in (True, False):
    eng = Engine(spec, synth.make_state_dict(spec, 0), use_graph=use_graph)
    sess = eng.session(B, H, W, False)
    for _ in range(3):
        sess.run(imgs)
    torch.cuda.synchronize()
    import time
    time.sleep(2.0)                     # let the GPU cool / clocks recover
    n = 40
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    evs[0].record()
    for i in range(n):
        sess.run(imgs)
        evs[i + 1].record()
    torch.cuda.synchronize()
    ts = [evs[i].elapsed_time(evs[i + 1]) for i in range(n)]
    print(f"graph={use_graph}: steps 0-4 {[round(t, 2) for t in ts[:5]]}  10-14 {[round(t, 2) for t in ts[10:15]]}  35-39 {[round(t, 2) for t in ts[35:]]}", flush=True)
    time.sleep(2.0)
    p = sess.profile(imgs)
    print(f"   profile pass after idle: sum {sum(p):.2f} ms", flush=True)
    for _ in range(20):
        sess.run(imgs)
    p = sess.profile(imgs)
    print(f"   profile pass right after 20 steps: sum {sum(p):.2f} ms", flush=True)
    del sess, eng
    torch.cuda.empty_cache()
