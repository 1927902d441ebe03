"""Host-side cost of HostPipeline.submit(): wall time per step over a long run next to the device-resident step, and the
time the host spends inside _issue / _finish / the enqueue between them (perf_counter around each). Run under gpurun:
    python tools/dev_pipeline_host_timing.py [--dets 10] [--steps 40]
"""
import argparse
from dataclasses import replace
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from densepose_torchscript_b200 import synth
from densepose_torchscript_b200.config import BUILTIN
from densepose_torchscript_b200.engine import Engine, HostPipeline


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dets", type=int, default=10)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--batch", type=int, default=8)
    a = ap.parse_args()
    spec = replace(BUILTIN["densepose_rcnn_R_50_FPN_s1x"], dets_per_image=a.dets)
    eng = Engine(spec, synth.make_state_dict(spec, 0))
    B, H, W = a.batch, 800, 1333
    host = torch.stack([synth.synthetic_image(H, W, seed=100 + i) for i in range(B)]).pin_memory()
    u8 = host.round().clamp(0, 255).to(torch.uint8).pin_memory()
    # device-resident step
    sess = eng.session(B, H, W, False)
    dev_in = host.cuda()
    for _ in range(3):
        sess.run(dev_in)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        sess.run(dev_in)
    e1.record()
    torch.cuda.synchronize()
    print(f"device-resident step {e0.elapsed_time(e1) / a.steps:.3f} ms")
    for extract in (False, True):
        src = u8 if extract else host
        pipe = HostPipeline(eng, B, H, W, extract, depth=2, extract=extract)
        acc = {"issue": 0.0, "finish": 0.0, "wait": 0.0}
        issue0, finish0 = pipe._issue, pipe._finish

        def issue(sl, _f=issue0):
            t0 = time.perf_counter()
            if sl["busy"]:
                sl["done"].synchronize()
            t1 = time.perf_counter()
            r = _f(sl)
            acc["wait"] += t1 - t0
            acc["issue"] += time.perf_counter() - t1
            return r

        def finish(t, _f=finish0):
            t0 = time.perf_counter()
            r = _f(t)
            acc["finish"] += time.perf_counter() - t0
            return r

        pipe._issue, pipe._finish = issue, finish
        for _ in range(6):
            pipe.submit(src)
        pipe.drain()
        torch.cuda.synchronize()
        for k in acc:
            acc[k] = 0.0
        t0 = time.perf_counter()
        for _ in range(a.steps):
            pipe.submit(src)
        t_sub = time.perf_counter() - t0
        pipe.drain()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        n = a.steps
        print(f"extract={extract}: wall {wall / n * 1e3:.3f} ms/step (submit loop only {t_sub / n * 1e3:.3f}); host per step: "
              f"wait-for-forward {acc['wait'] / n * 1e3:.3f}, issue {acc['issue'] / n * 1e3:.3f}, finish {acc['finish'] / n * 1e3:.3f}, "
              f"enqueue+rest {(t_sub - acc['wait'] - acc['issue'] - acc['finish']) / n * 1e3:.3f} ms; d2h {pipe.last_d2h_bytes} B")
        pipe.close()


if __name__ == "__main__":
    main()
