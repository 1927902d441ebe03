"""CUDA-event timing of the one-launch deconv (the four ConvTranspose2d phases as N blocks, conv_igemm phase_taps) at the
bench shape (800 ROIs x 28 x 28 x 512 -> 4 x 80 channels, fp32 channel-planar output) for the plan overrides the C-ABI
exposes: CTA pairs on / off, pipeline depth, chunks per barrier.
    python tools/dev_deconv_timing.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from densepose_torchscript_b200 import ops


def ev_time(fn, n=20):
    for _ in range(4):
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
    torch.manual_seed(0)
    R, P, Cin, Cp = 800, 28, 512, 80
    x = (torch.randn(R, P, P, Cin, device="cuda") * 0.5).to(torch.bfloat16)
    packed = (torch.randn(4 * Cp, 4 * Cin, device="cuda") * 0.02).to(torch.bfloat16)
    bias = torch.randn(4 * Cp, device="cuda")
    out = torch.empty(R, 4 * Cp, P, P, device="cuda")
    flops = 2.0 * R * P * P * 4 * Cp * 4 * Cin
    ref = None
    ms = ev_time(lambda: ops.conv2d(x, packed, bias, 2, 2, pad=1, planar=True, out=out, phase_taps=True, pair=1, tiled=True))
    print(f"tiled A loads (28 x 4 pixel boxes, 112-row tiles): {ms:.4f} ms")
    for pair in (1, 2):
        for stages in (0, 3, 4, 5, 6):
            for ks in (0, 1, 2):
                if pair == 2 and ks == 2:
                    continue
                try:
                    f = lambda: ops.conv2d(x, packed, bias, 2, 2, pad=1, planar=True, out=out, phase_taps=True, pair=pair,
                                           stages=stages, ks=ks)
                    ms = ev_time(f)
                    torch.cuda.synchronize()
                    if ref is None:
                        ref = out.clone()
                    same = bool(torch.equal(out, ref))
                    print(f"pair {pair} stages {stages} ks {ks}: {ms:.4f} ms  {flops / ms / 1e9:.0f} TFLOP/s  identical {same}")
                except Exception as ex:  # noqa: BLE001
                    print(f"pair {pair} stages {stages} ks {ks}: {str(ex)[:100]}")


if __name__ == "__main__":
    main()
