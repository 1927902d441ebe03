"""Where does the host-in / host-out pipeline lose time against the PCIe link? ms/step of HostPipeline for
several depths, next to the bare D2H copy of one step's outputs (no kernels running)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from densepose_torchscript_b200 import synth
from densepose_torchscript_b200.config import BUILTIN
from densepose_torchscript_b200.engine import Engine, HostPipeline

spec = BUILTIN["densepose_rcnn_R_50_FPN_s1x"]
eng = Engine(spec, synth.make_state_dict(spec, 0))
B, H, W = 8, 800, 1333
host = torch.stack([synth.synthetic_image(H, W, seed=100 + i) for i in range(B)]).contiguous().pin_memory()
K = int(os.environ.get("K", "12"))


def timed(pipe, k=K):
    for _ in range(3):
        pipe.submit(host)
    pipe.drain()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for sl in pipe.slots:
        sl["sess"].stream.wait_stream(torch.cuda.current_stream())
    for _ in range(k):
        pipe.submit(host)
    pipe.drain()
    for sl in pipe.slots:
        torch.cuda.current_stream().wait_stream(sl["sess"].stream)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k


for depth in (2, 3, 2, 3):
    pipe = HostPipeline(eng, B, H, W, False, depth=depth)
    ms = timed(pipe)
    print(f"depth {depth}: {ms:.2f} ms/step  {B / ms * 1e3:.1f} images/s  d2h {pipe.d2h_bytes / 1e6 / ms:.1f} GB/s", flush=True)
    if depth == 2:
        sl = pipe.slots[0]
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(4):
            for hbuf, dbuf in zip(sl["outs_host"], sl["outs_dev"]):
                hbuf.copy_(dbuf, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 4
        print(f"  bare D2H of one step's outputs: {ms:.2f} ms  {pipe.d2h_bytes / 1e6 / ms:.1f} GB/s", flush=True)
    pipe.close()
    del pipe
    torch.cuda.empty_cache()
