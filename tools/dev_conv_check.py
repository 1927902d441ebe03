"""Bring-up check of the tcgen05 implicit-GEMM conv on a real B200 (run under gpurun).

Each case runs the C-ABI conv on bf16-rounded inputs and compares with torch fp32 conv2d of the same
rounded inputs (TF32 off).  Cases run in child processes so a trapped kernel cannot poison the rest.
Usage: python tools/dev_conv_check.py            (driver: all cases, both A-operand modes)
       python tools/dev_conv_check.py --case K --tiled 0|1
"""
import argparse
import json
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

CASES = [
    # name, N, H, W, Cin, Cout, k, stride, pad, dil, relu, bias, res, res_shift, out_fp32, n_valid
    dict(name="1x1_tiny", N=1, H=16, W=16, Cin=64, Cout=64, k=1),
    dict(name="1x1_k256", N=1, H=25, W=42, Cin=256, Cout=128, k=1, bias=True),
    dict(name="3x3_tiny", N=1, H=16, W=16, Cin=64, Cout=64, k=3, pad=1),
    dict(name="3x3_odd", N=2, H=13, W=21, Cin=128, Cout=48, k=3, pad=1, bias=True),
    dict(name="3x3_256", N=1, H=50, W=84, Cin=256, Cout=256, k=3, pad=1, relu=True, bias=True),
    dict(name="1x1_res_512", N=2, H=25, W=42, Cin=128, Cout=512, k=1, relu=True, bias=True, res=True),
    dict(name="1x1_s2", N=2, H=50, W=84, Cin=256, Cout=128, k=1, stride=2, bias=True),
    dict(name="1x1_up2", N=1, H=50, W=84, Cin=512, Cout=256, k=1, bias=True, res=True, res_shift=1),
    dict(name="fc", N=300, H=1, W=1, Cin=1024, Cout=1024, k=1, relu=True, bias=True),
    dict(name="fc_small_out", N=1000, H=1, W=1, Cin=1024, Cout=6, k=1, bias=True, out_fp32=True),
    dict(name="roi_nvalid", N=9, H=28, W=28, Cin=256, Cout=512, k=3, pad=1, relu=True, bias=True, n_valid=5),
    dict(name="dil6", N=3, H=28, W=28, Cin=256, Cout=256, k=3, pad=6, dil=6),
    dict(name="dil12", N=3, H=28, W=28, Cin=256, Cout=256, k=3, pad=12, dil=12),
    dict(name="roi14", N=7, H=14, W=14, Cin=256, Cout=512, k=3, pad=1, relu=True, bias=True),
    dict(name="big_3x3", N=1, H=200, W=336, Cin=256, Cout=256, k=3, pad=1, relu=True, bias=True, time=True),
    dict(name="head_3x3", N=100, H=28, W=28, Cin=512, Cout=512, k=3, pad=1, relu=True, bias=True, time=True),
    dict(name="res4_1x1", N=8, H=50, W=84, Cin=1024, Cout=256, k=1, relu=True, bias=True, time=True),
    dict(name="res2_conv3", N=8, H=200, W=336, Cin=64, Cout=256, k=1, relu=True, bias=True, res=True, time=True),
    dict(name="res2_shortcut", N=8, H=200, W=336, Cin=64, Cout=256, k=1, bias=True, time=True),
    dict(name="res3_conv3", N=8, H=100, W=168, Cin=128, Cout=512, k=1, relu=True, bias=True, res=True, time=True),
    dict(name="res5_conv3", N=8, H=25, W=42, Cin=512, Cout=2048, k=1, relu=True, bias=True, res=True, time=True),
    dict(name="lateral2", N=8, H=200, W=336, Cin=256, Cout=256, k=1, bias=True, res=True, res_shift=1, time=True),
    dict(name="res2_conv2", N=8, H=200, W=336, Cin=64, Cout=64, k=3, pad=1, relu=True, bias=True, time=True),
    dict(name="m_tail", N=3, H=7, W=9, Cin=64, Cout=128, k=1, relu=True, bias=True, res=True),
    dict(name="planar_1x1", N=5, H=28, W=28, Cin=512, Cout=77, k=1, bias=True, planar=True, n_valid=4),
    dict(name="planar_3x3", N=2, H=14, W=14, Cin=128, Cout=90, k=3, pad=1, bias=True, planar=True),
]


def run_case(idx: int, mode: int):
    tiled = 1 if mode == 1 else 0
    epilogue = {0: 0, 1: 0, 2: 2, 3: 1}[mode]
    import torch
    import torch.nn.functional as F
    from densepose_torchscript_b200 import ops

    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    c = dict(stride=1, pad=0, dil=1, relu=False, bias=False, res=False, res_shift=0, out_fp32=False,
             n_valid=None, time=False, planar=False)
    c.update(CASES[idx])
    if mode == 2 and (c["out_fp32"] or c["planar"] or ((c["Cout"] + 15) // 16 * 16) % 64 != 0):
        print("RESULT " + json.dumps(dict(case=c["name"], tiled=mode, ok=True, skipped="not eligible for the staged epilogue")), flush=True)
        return
    g = torch.Generator(device="cpu").manual_seed(1234 + idx)
    dev = "cuda"
    x = torch.randn(c["N"], c["Cin"], c["H"], c["W"], generator=g).to(dev)
    w = (torch.randn(c["Cout"], c["Cin"], c["k"], c["k"], generator=g) / (c["Cin"] * c["k"] ** 2) ** 0.5).to(dev)
    b = torch.randn(c["Cout"], generator=g).to(dev) if c["bias"] else None
    xb = x.to(torch.bfloat16)
    wb = w.to(torch.bfloat16)
    ref = F.conv2d(xb.float(), wb.float(), b, stride=c["stride"], padding=c["pad"], dilation=c["dil"])
    Ho, Wo = ref.shape[2], ref.shape[3]
    res = None
    if c["res"]:
        if c["res_shift"]:
            rs = torch.randn(c["N"], c["Cout"], (Ho + 1) // 2, (Wo + 1) // 2, generator=g).to(dev).to(torch.bfloat16)
            ref = ref + F.interpolate(rs.float(), scale_factor=2, mode="nearest")[:, :, :Ho, :Wo]
        else:
            rs = torch.randn(c["N"], c["Cout"], Ho, Wo, generator=g).to(dev).to(torch.bfloat16)
            ref = ref + rs.float()
        res = rs.permute(0, 2, 3, 1).contiguous()
    if c["relu"]:
        ref = ref.relu()
    packed, bias_p, cin_pad, cout_pad = ops.pack_conv_weight(w, b)
    if res is not None and cout_pad != c["Cout"]:
        raise RuntimeError("test residual needs cout multiple of 16")
    x_nhwc = xb.permute(0, 2, 3, 1).contiguous()
    nv = None
    if c["n_valid"] is not None:
        nv = torch.tensor([c["n_valid"]], dtype=torch.int32, device=dev)
    shape = (c["N"], cout_pad, Ho, Wo) if c["planar"] else (c["N"], Ho, Wo, cout_pad)
    out = torch.full(shape, -777.0, device=dev,
                     dtype=torch.float32 if (c["out_fp32"] or c["planar"]) else torch.bfloat16)
    kw = dict(stride=c["stride"], pad=c["pad"], dil=c["dil"], relu=c["relu"], res=res, res_shift=c["res_shift"],
              out_fp32=c["out_fp32"], n_valid=nv, out=out, tiled=bool(tiled), epilogue=epilogue, planar=c["planar"])
    ops.conv2d(x_nhwc, packed, bias_p if c["bias"] else None, c["k"], c["k"], **kw)
    torch.cuda.synchronize()
    if c["planar"]:
        out = out.permute(0, 2, 3, 1)
    got = out[..., :c["Cout"]].permute(0, 3, 1, 2).float()
    nvv = c["n_valid"] if c["n_valid"] is not None else c["N"]
    err = (got[:nvv] - ref[:nvv]).abs().max().item()
    scale = ref[:nvv].abs().max().item()
    untouched = True
    if nvv < c["N"]:
        # the staged epilogue stores whole 128-row tiles: rows of invalid images inside the last valid tile
        # may be overwritten, everything after that tile must be untouched
        first = (nvv * Ho * Wo + 127) // 128 * 128 if mode != 1 else nvv * Ho * Wo
        flat = out.reshape(-1, out.shape[-1])
        untouched = bool((flat[first:] == -777.0).all().item())
    pad_ok = bool((out[:nvv, ..., c["Cout"]:] == 0).all().item()) if cout_pad > c["Cout"] and not c["bias"] else True
    tol = 2e-2 * max(scale, 1.0) if not (c["out_fp32"] or c["planar"]) else 2e-3 * max(scale, 1.0)
    r = dict(case=c["name"], tiled=mode, max_err=err, ref_max=scale, ok=bool(err < tol and untouched and pad_ok),
             untouched=untouched)
    if c["time"]:
        for _ in range(3):
            ops.conv2d(x_nhwc, packed, bias_p, c["k"], c["k"], **kw)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        iters = 20
        for _ in range(iters):
            ops.conv2d(x_nhwc, packed, bias_p, c["k"], c["k"], **kw)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        fl = 2.0 * c["N"] * Ho * Wo * c["Cout"] * c["Cin"] * c["k"] ** 2
        r["ms"] = ms
        r["tflops"] = fl / ms / 1e9
    print("RESULT " + json.dumps(r), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", type=int, default=-1)
    ap.add_argument("--tiled", type=int, default=0)
    ap.add_argument("--modes", default="0,1")
    a = ap.parse_args()
    if a.case >= 0:
        # child: run cases case..end in this process; a CUDA fault ends the process, the driver resumes after it
        for i in range(a.case, len(CASES)):
            print("BEGIN %d" % i, flush=True)
            run_case(i, a.tiled)
        return
    results = []
    for tiled in [int(m) for m in a.modes.split(",")]:
        start = 0
        while start < len(CASES):
            t0 = time.time()
            try:
                p = subprocess.run([sys.executable, os.path.abspath(__file__), "--case", str(start), "--tiled", str(tiled)],
                                   capture_output=True, text=True, timeout=420)
                out, err, rc = p.stdout, p.stderr, p.returncode
            except subprocess.TimeoutExpired as e:
                out = (e.stdout or b"").decode() if isinstance(e.stdout, bytes) else (e.stdout or "")
                err, rc = "TIMEOUT", -9
            last_begin = start - 1
            for l in out.splitlines():
                if l.startswith("BEGIN "):
                    last_begin = int(l[6:])
                if l.startswith("RESULT "):
                    r = json.loads(l[7:])
                    print(json.dumps(r), flush=True)
                    results.append(r)
            done = [r for r in results if r["tiled"] == tiled]
            if last_begin >= start and (len(done) == 0 or done[-1]["case"] != CASES[last_begin]["name"]):
                r = dict(case=CASES[last_begin]["name"], tiled=tiled, ok=False, rc=rc, tail=(out[-400:] + err[-1500:]))
                print(json.dumps(r), flush=True)
                results.append(r)
            start = max(last_begin, start) + 1
            print("# child wall %.1fs" % (time.time() - t0), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/dev_conv_check.json", "w") as f:
        json.dump(results, f, indent=1)
    print("PASS" if all(r.get("ok") for r in results) else "FAIL", sum(bool(r.get("ok")) for r in results), "/", len(results))


if __name__ == "__main__":
    main()
