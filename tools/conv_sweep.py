"""Times dpb200_conv2d on the conv shapes of the bench workload for a grid of (block_n, stages) overrides.
Run under gpurun; prints one line per (shape, config): median ms of 7 launches, TFLOP/s, GB/s of algorithmic bytes.
Used to choose the plan heuristics in conv_plan_build (conv_igemm.cu)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

from densepose_torchscript_b200 import ops

# name, N, H, W, Cin, Cout, k, stride, pad, residual
SHAPES = [
    ("res2.conv1 1x1 256->64", 8, 200, 336, 256, 64, 1, 1, 0, False),
    ("res2.conv2 3x3 64->64", 8, 200, 336, 64, 64, 3, 1, 1, False),
    ("res2.conv3 1x1 64->256 +res", 8, 200, 336, 64, 256, 1, 1, 0, True),
    ("res3.conv2 3x3 128->128", 8, 100, 168, 128, 128, 3, 1, 1, False),
    ("res3.conv3 1x1 128->512 +res", 8, 100, 168, 128, 512, 1, 1, 0, True),
    ("res4.conv1 1x1 1024->256", 8, 50, 84, 1024, 256, 1, 1, 0, False),
    ("res4.conv2 3x3 256->256", 8, 50, 84, 256, 256, 3, 1, 1, False),
    ("res4.conv3 1x1 256->1024 +res", 8, 50, 84, 256, 1024, 1, 1, 0, True),
    ("res5.conv2 3x3 512->512", 8, 25, 42, 512, 512, 3, 1, 1, False),
    ("fpn_out2 3x3 256->256", 8, 200, 336, 256, 256, 3, 1, 1, False),
    ("head 3x3 512->512 (200 rois)", 200, 28, 28, 512, 512, 3, 1, 1, False),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    ap.add_argument("--tiled", action="store_true", help="A operand through tiled (non-im2col) TMA boxes")
    a = ap.parse_args()
    torch.manual_seed(0)
    for name, n, h, w, cin, cout, k, stride, pad, has_res in SHAPES:
        if a.only and a.only not in name:
            continue
        x = (torch.randn(n, h, w, cin, device="cuda") * 0.5).to(torch.bfloat16)
        wt = torch.randn(cout, cin, k, k) * (1.0 / (cin * k * k) ** 0.5)
        packed, bias, _, _ = ops.pack_conv_weight(wt, torch.randn(cout))
        packed, bias = packed.cuda(), bias.cuda()
        ho = (h + 2 * pad - (k - 1) - 1) // stride + 1
        wo = (w + 2 * pad - (k - 1) - 1) // stride + 1
        res = (torch.randn(n, ho, wo, cout, device="cuda") * 0.5).to(torch.bfloat16) if has_res else None
        out = torch.empty(n, ho, wo, packed.shape[0], device="cuda", dtype=torch.bfloat16)
        flops = 2.0 * n * ho * wo * cout * cin * k * k
        by = x.numel() * 2 + out.numel() * 2 + packed.numel() * 2 + (res.numel() * 2 if has_res else 0)
        cfgs = [(0, 0, 0)]
        for bn in (64, 128, 256):
            if bn <= packed.shape[0] and packed.shape[0] % bn == 0:
                for ks in (1, 2):
                    for st in (0, 2, 3, 4, 6):
                        cfgs.append((bn, st, ks))
        best = None
        for bn, st, ks in cfgs:
            ts = []
            try:
                for _ in range(9):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    ops.conv2d(x, packed, bias, k, k, stride, pad, 1, True, res=res, out=out, block_n=bn, stages=st,
                               tiled=a.tiled, ks=ks)
                    e1.record()
                    torch.cuda.synchronize()
                    ts.append(e0.elapsed_time(e1))
            except Exception as ex:  # noqa: BLE001
                print(f"{name:32s} block_n {bn:3d} stages {st} ks {ks}  -> {str(ex)[:80]}")
                continue
            ts = sorted(ts[2:])
            ms = ts[len(ts) // 2]
            tag = ""
            if best is None or ms < best[0]:
                best = (ms, bn, st, ks)
            print(f"{name:32s} block_n {bn:3d} stages {st} ks {ks}  {ms:8.4f} ms  {flops / ms / 1e9:8.1f} TFLOP/s  {by / ms / 1e6:8.1f} GB/s{tag}")
        print(f"{name:32s} BEST block_n {best[1]} stages {best[2]} ks {best[3]} {best[0]:.4f} ms\n")


if __name__ == "__main__":
    main()
