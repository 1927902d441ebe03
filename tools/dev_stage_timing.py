"""CUDA-event timing of the HBM-bound stage kernels at the bench shapes (batch 8, 800x1344 padded):
    python tools/dev_stage_timing.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from densepose_torchscript_b200 import ops


def ev_time(fn, n=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


g = torch.Generator(device="cuda").manual_seed(0)
B = 8
a = torch.randn(B, 200, 336, 256, device="cuda", generator=g).to(torch.bfloat16)
bs = [torch.randn(B, 100, 168, 256, device="cuda", generator=g).to(torch.bfloat16) for _ in range(3)]
print("decoder_merge  ms", round(ev_time(lambda: ops.decoder_merge(a, *bs)), 4))
x = torch.randn(B, 50, 84, 256, device="cuda", generator=g).to(torch.bfloat16)
print("upsample2x p4  ms", round(ev_time(lambda: ops.upsample2x(x)), 4))
x5 = torch.randn(B, 25, 42, 256, device="cuda", generator=g).to(torch.bfloat16)
print("upsample2x p5  ms", round(ev_time(lambda: ops.upsample2x(x5)), 4))
stem = torch.randn(B, 400, 672, 64, device="cuda", generator=g).to(torch.bfloat16)
print("maxpool        ms", round(ev_time(lambda: ops.maxpool3x3s2(stem)), 4))
img = torch.rand(B, 800, 1333, 3, device="cuda", generator=g) * 255
print("preprocess f32 ms", round(ev_time(lambda: ops.preprocess(img, 1.0, (103.53, 116.28, 123.675), (1., 1., 1.))), 4))
u8 = img.to(torch.uint8)
print("preprocess u8  ms", round(ev_time(lambda: ops.preprocess(u8, 1.0, (103.53, 116.28, 123.675), (1., 1., 1.))), 4))
