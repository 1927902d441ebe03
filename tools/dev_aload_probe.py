"""Is a small-N conv bound by its A-operand (im2col TMA) loads rather than by the MMAs? The same 2x2-tap conv over the
bench's head output (800 x 28 x 28 x 512 -> 320 channels) with the N tile at 80 / 160 (/ 64 / 128 / 256 over 256 channels):
FLOPs are identical, the number of A tile loads halves each time the N tile doubles.
    python tools/dev_aload_probe.py
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
    R, P, Cin = 800, 28, 512
    x = (torch.randn(R, P, P, Cin, device="cuda") * 0.5).to(torch.bfloat16)
    for cout, bns in ((320, (80, 160)), (256, (64, 128, 256))):
        packed = (torch.randn(cout, 4 * Cin, device="cuda") * 0.02).to(torch.bfloat16)
        bias = torch.randn(cout, device="cuda")
        for fp32 in (True, False):
            out = torch.empty(R, P + 1, P + 1, cout, device="cuda", dtype=torch.float32 if fp32 else torch.bfloat16)
            flops = 2.0 * R * (P + 1) * (P + 1) * cout * 4 * Cin
            for bn in bns:
                for pair, tiled in ((1, False), (2, False), (1, True)):
                    try:
                        ms = ev_time(lambda: ops.conv2d(x, packed, bias, 2, 2, pad=1, out=out, block_n=bn, pair=pair, tiled=tiled))
                        print(f"cout {cout} fp32_out {fp32} block_n {bn:3d} pair {pair} tiled {tiled}: {ms:.4f} ms  {flops / ms / 1e9:.0f} TFLOP/s")
                    except Exception as ex:  # noqa: BLE001
                        print(f"cout {cout} block_n {bn} pair {pair}: {str(ex)[:90]}")


if __name__ == "__main__":
    main()
