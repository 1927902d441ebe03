"""Stage-isolated parity: every CUDA kernel, called through the C-ABI, against the CPU oracle on identical
(seeded) inputs.  Integer / index / ordering results must be bit-exact; floating-point tolerances are
written beside each check.  Run on a B200:  python -m pytest tests -m gpu -x -q
"""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from _util import bf16, frac_equal, nchw, nhwc_bf16_cuda, rel_l2
from oracle import aten_interp as AI
from oracle import densepose_oracle as O

# the host's ATen build is bit-identical to the restatements in oracle/aten_interp.py on x86 (tests/test_oracle.py);
# where that holds the kernels are also compared with F.interpolate itself
X86 = torch.backends.cpu.get_cpu_capability() in ("AVX2", "AVX512")

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from densepose_torchscript_b200 import ops as _ops
    from densepose_torchscript_b200 import _lib
    _lib.require_device()
    return _ops


# ----------------------------------------------------------------------------------------------- conv
CONV_CASES = [
    # N, H, W, Cin, Cout, k, stride, pad, dil, relu, res(0 none,1 same,2 up2)
    (1, 16, 16, 64, 64, 1, 1, 0, 1, False, 0),
    (2, 13, 21, 128, 48, 3, 1, 1, 1, False, 0),
    (1, 50, 84, 256, 256, 3, 1, 1, 1, True, 0),
    (2, 25, 42, 128, 512, 1, 1, 0, 1, True, 1),
    (2, 50, 84, 256, 128, 1, 2, 0, 1, False, 0),
    (1, 50, 84, 512, 256, 1, 1, 0, 1, False, 2),
    (300, 1, 1, 1024, 1024, 1, 1, 0, 1, True, 0),
    (3, 28, 28, 256, 256, 3, 1, 6, 6, False, 0),
    (3, 28, 28, 256, 256, 3, 1, 12, 12, False, 0),
    (7, 14, 14, 256, 512, 3, 1, 1, 1, True, 0),
]


@pytest.mark.parametrize("case", CONV_CASES)
@pytest.mark.parametrize("tiled", [False, True])
def test_conv_igemm_matches_oracle(ops, case, tiled):
    N, H, W, Cin, Cout, k, stride, pad, dil, relu, res_mode = case
    g = torch.Generator().manual_seed(hash(case) % 100000)
    x = bf16(torch.randn(N, Cin, H, W, generator=g))
    w = bf16(torch.randn(Cout, Cin, k, k, generator=g) / math.sqrt(Cin * k * k))
    b = torch.randn(Cout, generator=g)
    ref = F.conv2d(x, w, b, stride=stride, padding=pad, dilation=dil)        # wrappers.py:105-107 on the same bf16 operands
    res = None
    if res_mode == 1:
        res = bf16(torch.randn_like(ref))
        ref = ref + res
    elif res_mode == 2:
        res = bf16(torch.randn(N, Cout, (ref.shape[2] + 1) // 2, (ref.shape[3] + 1) // 2, generator=g))
        ref = ref + F.interpolate(res, scale_factor=2.0, mode="nearest")[:, :, :ref.shape[2], :ref.shape[3]]   # fpn.py:152-154
    if relu:
        ref = F.relu(ref)
    packed, bias, _, cout_pad = ops.pack_conv_weight(w.cuda(), b.cuda())
    out = ops.conv2d(nhwc_bf16_cuda(x), packed, bias, k, k, stride=stride, pad=pad, dil=dil, relu=relu,
                     res=None if res is None else nhwc_bf16_cuda(res), res_shift=1 if res_mode == 2 else 0,
                     tiled=tiled)
    torch.cuda.synchronize()
    got = nchw(out[..., :Cout])
    # bf16 output: half an ulp of the largest magnitude (2^-9 relative) plus fp32 accumulation-order noise
    tol = float(ref.abs().max()) * 2.0 ** -8 + 1e-3
    assert float((got - ref).abs().max()) <= tol
    assert rel_l2(got, bf16(ref)) < 3e-3


PAIR_CASES = [
    # N, H, W, Cin, Cout, k, pad, dil, relu, res(0 none,1 same,2 up2), block_n, n_valid
    (6, 28, 28, 128, 512, 3, 1, 1, True, 0, 256, 0),       # head-conv shape, 37 tiles (odd: last pair half empty)
    (6, 28, 28, 128, 512, 3, 1, 1, True, 0, 256, 3),       # device-side count: tiles past the valid rows
    (2, 50, 84, 256, 256, 3, 1, 1, False, 0, 256, 0),      # FPN output / RPN conv shape
    (2, 25, 42, 64, 256, 1, 0, 1, True, 1, 256, 0),        # residual through TMA
    (1, 50, 84, 128, 256, 1, 0, 1, False, 2, 128, 0),      # gathered top-down residual, two N blocks of 128
    (3, 28, 28, 256, 64, 3, 6, 6, False, 0, 64, 0),        # dilated, narrow tile
]


@pytest.mark.parametrize("case", PAIR_CASES)
def test_conv_cta_pairs_match_single_cta_bitwise(ops, case):
    """cta_group::2 launch (clusters of two CTAs, weights split over the pair) == the single-CTA kernel, bit for bit,
    and both within tolerance of F.conv2d."""
    N, H, W, Cin, Cout, k, pad, dil, relu, res_mode, block_n, n_valid = case
    g = torch.Generator().manual_seed(sum(case))
    x = bf16(torch.randn(N, Cin, H, W, generator=g))
    w = bf16(torch.randn(Cout, Cin, k, k, generator=g) / math.sqrt(Cin * k * k))
    b = torch.randn(Cout, generator=g)
    ref = F.conv2d(x, w, b, padding=pad, dilation=dil)
    res = None
    if res_mode == 1:
        res = bf16(torch.randn_like(ref))
        ref = ref + res
    elif res_mode == 2:
        res = bf16(torch.randn(N, Cout, (H + 1) // 2, (W + 1) // 2, generator=g))
        ref = ref + F.interpolate(res, scale_factor=2.0, mode="nearest")[:, :, :H, :W]
    if relu:
        ref = F.relu(ref)
    packed, bias, _, cout_pad = ops.pack_conv_weight(w.cuda(), b.cuda())
    nv = torch.tensor([n_valid], dtype=torch.int32, device="cuda") if n_valid else None
    outs = []
    for pair in (1, 2):
        out = torch.full((N, H, W, cout_pad), -5.0, device="cuda", dtype=torch.bfloat16)
        ops.conv2d(nhwc_bf16_cuda(x), packed, bias, k, k, pad=pad, dil=dil, relu=relu,
                   res=None if res is None else nhwc_bf16_cuda(res), res_shift=1 if res_mode == 2 else 0,
                   block_n=block_n, epilogue=2, pair=pair, n_valid=nv, out=out)
        torch.cuda.synchronize()
        outs.append(out)
    nvis = n_valid if n_valid else N
    assert torch.equal(outs[0][:nvis], outs[1][:nvis])
    got = nchw(outs[1][:nvis, ..., :Cout])
    tol = float(ref.abs().max()) * 2.0 ** -8 + 1e-3
    assert float((got - ref[:nvis]).abs().max()) <= tol
    if n_valid:
        # rows of images past the count may be touched only inside the last (partial) 256-row pair tile
        flat = outs[1].view(-1, cout_pad)
        first_free = ((n_valid * H * W + 255) // 256) * 256
        assert bool((flat[first_free:] == -5.0).all())


@pytest.mark.parametrize("tiled", [False, True, "pair"])
def test_deconv_four_phases_in_one_launch(ops, tiled):
    """ConvTranspose2d(k=4,s=2,p=1) (chart.py:45-59) as one GEMM launch whose N blocks are the output-parity phases."""
    from densepose_torchscript_b200.weights import _pack_khwc
    g = torch.Generator().manual_seed(11)
    R, P, Cin, Cout = 5, 14, 128, 77
    x = bf16(torch.randn(R, Cin, P, P, generator=g))
    wt = bf16(torch.randn(Cin, Cout, 4, 4, generator=g) / math.sqrt(Cin * 4))
    bt = torch.randn(Cout, generator=g)
    ref = F.conv_transpose2d(x, wt, bt, stride=2, padding=1)                   # [R, Cout, 2P, 2P]
    ph = []
    for py in range(2):
        for px in range(2):
            kys = [3 - 2 * t for t in range(2)] if py == 0 else [2 - 2 * t for t in range(2)]
            kxs = [3 - 2 * t for t in range(2)] if px == 0 else [2 - 2 * t for t in range(2)]
            ph.append(_pack_khwc(wt[:, :, kys, :][:, :, :, kxs].permute(1, 2, 3, 0), bt, "cuda"))
    packed = torch.cat([q[0] for q in ph], 0).contiguous()
    bias = torch.cat([q[1] for q in ph], 0).contiguous()
    cp = ph[0][3]
    nv = torch.tensor([4], dtype=torch.int32, device="cuda")
    out = torch.full((R, 4 * cp, P, P), -3.0, device="cuda")
    ops.conv2d(nhwc_bf16_cuda(x), packed, bias, 2, 2, pad=1, planar=True, out=out, phase_taps=True, tiled=tiled is True,
               pair=2 if tiled == "pair" else 1, n_valid=nv)       # "pair": CTA pairs with the direct (fp32, planar) epilogue
    torch.cuda.synchronize()
    low = out.view(R, 2, 2, cp, P, P)[:, :, :, :Cout].cpu()                  # [r, py, px, c, y, x]
    got = low.permute(0, 3, 4, 1, 5, 2).reshape(R, Cout, 2 * P, 2 * P)          # (2y+py, 2x+px)
    assert float((got[:4] - ref[:4]).abs().max()) < 2e-4                        # fp32 accumulate, fp32 store
    assert bool((out[4:] == -3.0).all())


STRICT_CASES = [
    # N, H, W, Cin, Cout, k, stride, pad, dil, relu, res(0 none, 1 same, 2 up2), fp32 out
    (2, 25, 42, 64, 64, 1, 1, 0, 1, True, 0, False),
    (1, 50, 84, 256, 256, 3, 1, 1, 1, True, 0, False),
    (2, 25, 42, 128, 512, 1, 1, 0, 1, True, 1, False),
    (2, 50, 84, 256, 128, 1, 2, 0, 1, False, 0, False),
    (1, 50, 84, 512, 256, 1, 1, 0, 1, False, 2, False),
    (40, 1, 1, 12544, 1024, 1, 1, 0, 1, True, 0, False),      # FC1: K = 3 x 12544
    (3, 28, 28, 512, 512, 3, 1, 1, 1, True, 0, False),        # head conv
    (2, 50, 84, 256, 15, 1, 1, 0, 1, False, 0, True),         # RPN predictor: fp32 output
    (3, 28, 28, 256, 256, 3, 1, 12, 12, False, 0, False),     # ASPP, dilation 12
]


@pytest.mark.parametrize("case", STRICT_CASES)
def test_conv_strict_mode_is_fp32_class(ops, case):
    """Strict numerics: fp32 operands as bf16 (hi, lo) pairs, x_hi*w_hi + x_hi*w_lo + x_lo*w_hi accumulated in fp32 by
    three tcgen05.mma passes. Against F.conv2d in fp64 on the un-rounded fp32 operands the error must be that of 16-bit
    operands (2^-16-class), ~100x below the bf16 path's, and the returned (hi, lo) pair must be the split of one fp32."""
    N, H, W, Cin, Cout, k, stride, pad, dil, relu, res_mode, fp32_out = case
    g = torch.Generator().manual_seed(sum(case[:9]))
    x = torch.randn(N, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, k, k, generator=g) / math.sqrt(Cin * k * k)
    b = torch.randn(Cout, generator=g)
    ref = F.conv2d(x.double(), w.double(), b.double(), stride=stride, padding=pad, dilation=dil)
    res = None
    if res_mode == 1:
        res = torch.randn(ref.shape, generator=g)
        ref = ref + res.double()
    elif res_mode == 2:
        res = torch.randn(N, Cout, (ref.shape[2] + 1) // 2, (ref.shape[3] + 1) // 2, generator=g)
        ref = ref + F.interpolate(res.double(), scale_factor=2.0, mode="nearest")[:, :, :ref.shape[2], :ref.shape[3]]
    if relu:
        ref = F.relu(ref)
    to_nhwc = lambda t: t.permute(0, 2, 3, 1).contiguous().cuda()      # noqa: E731
    xh, xl = ops.split_bf16(to_nhwc(x))
    rh, rl = ops.split_bf16(to_nhwc(res)) if res is not None else (None, None)
    packed, bias, _, cout_pad = ops.pack_conv_weight(w.cuda(), b.cuda(), strict=True)
    out = ops.conv2d(xh, packed, bias, k, k, stride=stride, pad=pad, dil=dil, relu=relu, res=rh, res_lo=rl,
                     res_shift=1 if res_mode == 2 else 0, x_lo=xl, out_fp32=fp32_out)
    torch.cuda.synchronize()
    if fp32_out:
        got = nchw(out[..., :Cout]).double()
    else:
        hi, lo = out
        val = hi.float() + lo.float()
        assert bool((lo.float().abs() <= hi.float().abs() * 2.0 ** -8 + 1e-30).all())     # lo = bf16(x - bf16(x)): below half an ulp of hi
        got = nchw(val[..., :Cout]).double()
    err = float((got - ref).abs().max())
    scale = float(ref.abs().max())
    assert err <= scale * 2.0 ** -13, (err, scale)                      # bf16 path: ~2^-8 of the largest magnitude
    # bf16 path: ~3e-3. Measured here: 0.6-1.5e-5 for K <= 3 x 4608 and 4.6e-5 for FC1's K = 3 x 12544 — the tensor
    # core's fp32 accumulator truncates when it aligns addends, a bias that grows with the number of K steps
    assert float((got - ref).norm() / ref.norm()) < (2e-5 if Cin * k * k <= 4608 else 1e-4)


@pytest.mark.parametrize("case", [(1, 25, 42, 512, 512, 3, 1), (1, 50, 84, 1024, 256, 1, 0), (1, 50, 84, 256, 1024, 1, 0),
                                  (1000, 1, 1, 12544, 1024, 1, 0)])
def test_conv_result_does_not_depend_on_the_n_tile(ops, case):
    """Few-tile launches (batch 1) pick a narrower N tile so that more SMs work (conv_plan_build); the K order of every
    output element is unchanged, so the automatic plan must equal the forced 256- and 64-wide ones bit for bit."""
    N, H, W, Cin, Cout, k, pad = case
    g = torch.Generator().manual_seed(5)
    x = (torch.randn(N, H, W, Cin, generator=g) * 0.5).to(torch.bfloat16).cuda()
    w = torch.randn(Cout, Cin, k, k, generator=g) / math.sqrt(Cin * k * k)
    packed, bias, _, _ = ops.pack_conv_weight(w, torch.randn(Cout, generator=g))
    packed, bias = packed.cuda(), bias.cuda()
    res = (torch.randn(N, H, W, Cout, generator=g) * 0.5).to(torch.bfloat16).cuda()
    outs = [ops.conv2d(x, packed, bias, k, k, 1, pad, 1, True, res=res, block_n=bn) for bn in (0, 256, 64)]
    torch.cuda.synchronize()
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    ref = F.conv2d(x.float().permute(0, 3, 1, 2).cpu(), bf16(w), bias.cpu(), padding=pad).permute(0, 2, 3, 1) + res.float().cpu()
    err = (outs[0].float().cpu() - ref.clamp_min(0)).abs().max()
    assert float(err) <= float(ref.abs().max()) * 2.0 ** -7


def test_conv_fp32_out_and_n_valid(ops):
    g = torch.Generator().manual_seed(5)
    x = bf16(torch.randn(9, 256, 28, 28, generator=g))
    w = bf16(torch.randn(6, 256, 1, 1, generator=g) / 16)
    b = torch.randn(6, generator=g)
    packed, bias, _, cout_pad = ops.pack_conv_weight(w.cuda(), b.cuda())
    out = torch.full((9, 28, 28, cout_pad), -7.0, device="cuda")
    nv = torch.tensor([5], dtype=torch.int32, device="cuda")
    ops.conv2d(nhwc_bf16_cuda(x), packed, bias, 1, 1, out_fp32=True, n_valid=nv, out=out)
    torch.cuda.synchronize()
    ref = F.conv2d(x, w, b)
    assert float((nchw(out[:5, ..., :6]) - ref[:5]).abs().max()) < 2e-4      # fp32 accumulate, fp32 store
    assert bool((out[5:] == -7.0).all())                                      # images >= n_valid untouched
    assert bool((out[:5, ..., 6:] == 0).all())                                # padded output channels are zero


# ----------------------------------------------------------------------------------------------- preprocess
def _stem_image(ops, dst, wp):
    full = ops.stem_to_image(dst)[0]                 # [Hp, Wp + 8, 4], padded-image column x at x + 4
    return full, full[:, 4:4 + wp, :3].permute(2, 0, 1).float().cpu()


@pytest.mark.parametrize("h0,w0", [(240, 600), (300, 200), (97, 131), (1080, 1920)])
@pytest.mark.parametrize("variant", [0, 1])
def test_preprocess_float_is_bit_exact(ops, h0, w0, variant):
    """defaults.py:85-89 + rcnn.py:156-181 on a float image: the resized fp32 pixels are ATen's to the last bit
    (variant 0: multi-threaded reference = separable kernel; 1: single-threaded = channels-last kernel), so the
    normalised bf16 stem input is identical, not merely close."""
    spec = O.SPECS["densepose_rcnn_R_50_FPN_s1x"]
    g = torch.Generator().manual_seed(h0)
    img = torch.rand(h0, w0, 3, generator=g) * 255.0
    k = O.resize_scale(h0, w0, spec)
    hr, wr = int(math.floor(h0 * k)), int(math.floor(w0 * k))
    x = img.permute(2, 0, 1)[None].contiguous().numpy()
    resized = torch.from_numpy(AI.upsample_bilinear_f32(x, hr, wr, k, k, kernel="separable" if variant == 0 else "channels_last"))[0]
    if X86 and variant == 0 and torch.get_num_threads() > 1:
        image, _, _ = O.predictor_resize(img, spec)                              # F.interpolate itself
        assert torch.equal(image, resized)
    ref, padding = O.preprocess_image(resized, spec, O.Numerics("bf16"))          # [1,3,Hp,Wp]
    dst, (hr_, wr_, hp, wp) = ops.preprocess(img[None].cuda().contiguous(), k, spec.pixel_mean, spec.pixel_std, variant=variant)
    torch.cuda.synchronize()
    assert (hr_, wr_) == (hr, wr) and (hp, wp) == tuple(ref.shape[2:])
    assert tuple(dst.shape) == (1, hp // 2, wp // 2 + 4, 16)
    full, got = _stem_image(ops, dst, wp)
    assert torch.equal(got, ref[0])
    assert bool((full[:, :4] == 0).all()) and bool((full[:, 4 + wp:] == 0).all()) and bool((full[..., 3] == 0).all())
    assert bool((full[hr:] == 0).all()) and bool((full[:, 4 + wr:] == 0).all())   # zero padding in normalised space


@pytest.mark.parametrize("h0,w0", [(240, 600), (300, 200), (97, 131), (1080, 1920), (800, 1333), (64, 96)])
def test_preprocess_uint8_is_aten_fixed_point(ops, h0, w0):
    """run.py:33-36 feeds torch.from_numpy(cv2 image): the reference resizes in uint8 (ATen's int16 fixed-point two-pass
    bilinear, oracle/aten_interp.py). The kernel must return exactly those pixels (round 1 rounded a float bilinear:
    +-1 on 11-20 % of them)."""
    spec = O.SPECS["densepose_rcnn_R_50_FPN_s1x"]
    g = torch.Generator().manual_seed(w0)
    img = torch.randint(0, 256, (h0, w0, 3), dtype=torch.uint8, generator=g)
    k = O.resize_scale(h0, w0, spec)
    resized = torch.from_numpy(AI.upsample_bilinear_u8(img.numpy(), k)).permute(2, 0, 1)      # uint8 CHW
    if X86:
        image, _, _ = O.predictor_resize(img, spec)
        assert image.dtype == torch.uint8 and torch.equal(image, resized)
    ref, _ = O.preprocess_image(resized, spec, O.Numerics("bf16"))
    dst, (hr, wr, hp, wp) = ops.preprocess(img[None].cuda().contiguous(), k, spec.pixel_mean, spec.pixel_std)
    torch.cuda.synchronize()
    full, got = _stem_image(ops, dst, wp)
    assert torch.equal(got, ref[0])
    assert bool((full[hr:] == 0).all()) and bool((full[:, 4 + wr:] == 0).all())


def test_maxpool_exact(ops):
    g = torch.Generator().manual_seed(1)
    x = bf16(torch.randn(2, 64, 38, 50, generator=g)).relu()
    ref = F.max_pool2d(x, 3, 2, 1)                                              # resnet.py:353
    got = nchw(ops.maxpool3x3s2(nhwc_bf16_cuda(x)))
    assert torch.equal(got, ref)


def test_upsample_and_decoder_merge(ops):
    g = torch.Generator().manual_seed(2)
    x = bf16(torch.randn(2, 256, 13, 21, generator=g))
    ref = bf16(F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False))   # roi_head.py:63
    got = nchw(ops.upsample2x(nhwc_bf16_cuda(x)))
    assert frac_equal(got, ref) > 0.99 and rel_l2(got, ref) < 2e-3
    a = bf16(torch.randn(2, 256, 26, 42, generator=g))
    bs = [bf16(torch.randn(2, 256, 13, 21, generator=g)) for _ in range(3)]
    ref = a
    for t in bs:
        ref = ref + F.interpolate(t, scale_factor=2, mode="bilinear", align_corners=False)   # roi_head.py:73-77
    ref = bf16(ref)
    got = nchw(ops.decoder_merge(nhwc_bf16_cuda(a), *[nhwc_bf16_cuda(t) for t in bs]))
    assert frac_equal(got, ref) > 0.98 and rel_l2(got, ref) < 3e-3


def _up2_exact(x):
    """The kernels' x2 bilinear arithmetic, emulated exactly: top = fma(lx, q01, hx*q00), bot likewise, out =
    fma(ly, bot, hy*top) in fp32. With bf16 inputs and weights in {0, .25, .75, 1} every product has <= 10 significant
    bits and every fp32 rounding can be reproduced by rounding the float64 value."""
    n, c, h, w = x.shape
    xd = x.double()

    def taps(size):
        d = torch.arange(2 * size, dtype=torch.float64)
        real = (0.5 * (d + 0.5) - 0.5).clamp(min=0)
        i0 = real.floor().long()
        i1 = torch.where(i0 < size - 1, i0 + 1, i0)
        l = real - i0
        return i0, i1, l

    y0, y1, ly = taps(h)
    x0, x1, lx = taps(w)
    f32 = lambda t: t.float().double()                                   # noqa: E731  one fp32 rounding
    hx, hy = 1 - lx, 1 - ly
    q00, q01 = xd[:, :, y0][:, :, :, x0], xd[:, :, y0][:, :, :, x1]
    q10, q11 = xd[:, :, y1][:, :, :, x0], xd[:, :, y1][:, :, :, x1]
    top = f32(lx * q01 + f32(hx * q00))
    bot = f32(lx * q11 + f32(hx * q10))
    return f32(ly[:, None] * bot + f32(hy[:, None] * top))


def test_upsample_and_decoder_merge_exact_arithmetic(ops):
    """The 2x2-block kernels (9 loads per 4 outputs) must give, bit for bit, the per-output expression above: upsample =
    bf16(0 + lerp), merge = bf16(((a + up(b3)) + up(b4)) + up(b5)) with fp32 adds (roi_head.py:73-77)."""
    g = torch.Generator().manual_seed(7)
    for (h, w) in ((13, 21), (1, 5), (4, 1), (25, 42)):
        x = bf16(torch.randn(2, 256, h, w, generator=g))
        want = _up2_exact(x).float().to(torch.bfloat16)
        got = ops.upsample2x(nhwc_bf16_cuda(x)).permute(0, 3, 1, 2).cpu()
        assert torch.equal(got, want), (h, w)
    a = bf16(torch.randn(2, 256, 26, 42, generator=g))
    bs = [bf16(torch.randn(2, 256, 13, 21, generator=g)) for _ in range(3)]
    acc = a.double()
    for t in bs:
        acc = (acc + _up2_exact(t)).float().double()
    got = ops.decoder_merge(nhwc_bf16_cuda(a), *[nhwc_bf16_cuda(t) for t in bs]).permute(0, 3, 1, 2).cpu()
    assert torch.equal(got, acc.float().to(torch.bfloat16))


# ----------------------------------------------------------------------------------------------- RPN
def _rpn_inputs(seed, sizes, quant=None):
    g = torch.Generator().manual_seed(seed)
    logits, deltas, heads = [], [], []
    for (h, w) in sizes:
        lg = torch.randn(1, 3, h, w, generator=g) * 2.0
        if quant:
            lg = torch.round(lg / quant) * quant
        dl = torch.randn(1, 12, h, w, generator=g) * 0.5
        dl[:, 2::4] += 1.0
        dl[:, 3::4] += 1.0
        logits.append(lg); deltas.append(dl)
        head = torch.zeros(1, h, w, 16)
        head[..., :3] = lg.permute(0, 2, 3, 1)
        head[..., 3:15] = dl.permute(0, 2, 3, 1)
        heads.append(head.cuda().contiguous())
    return logits, deltas, heads


def test_rpn_proposals_match_oracle(ops):
    spec = O.SPECS["densepose_rcnn_R_50_FPN_s1x"]
    sizes = [(64, 96), (32, 48), (16, 24), (8, 12), (4, 6)]            # a 256x384 padded image
    Hp, Wp = 256, 384
    logits, deltas, heads = _rpn_inputs(11, sizes)
    lg, dl = O.rpn_flatten(logits, deltas)
    anchors = O.grid_anchors(sizes)
    props = [O.apply_deltas(d.reshape(-1, 4), a, (1.0, 1.0, 1.0, 1.0)).view(1, -1, 4) for a, d in zip(anchors, dl)]
    ref = O.find_top_rpn_proposals(props, lg, (Wp, Hp), spec)
    boxes, scores, counts, dbg = ops.rpn_proposals(heads, clip_x=float(Hp), clip_y=float(Wp))
    torch.cuda.synchronize()
    # per-level top-k: selected scores are bit-exact and in descending order; decoded+clipped boxes within fp32 exp() ulps
    for l, (p, s) in enumerate(zip(props, lg)):
        k = min(1000, s.shape[1])
        ts, ti = s[0].topk(k)
        assert int(dbg["cand_count"][0, l]) == k
        assert torch.equal(dbg["cand_scores"][0, l, :k].cpu(), ts)
        ref_boxes = O.clip_boxes(p[0, ti], (Wp, Hp))                    # quirk 1: swapped extents
        assert torch.allclose(dbg["cand_boxes"][0, l, :k].cpu(), ref_boxes, rtol=1e-5, atol=1e-3)
    n = int(counts[0])
    assert n == len(ref["proposal_boxes"])
    assert torch.equal(scores[0, :n].cpu(), ref["objectness_logits"])   # NMS keep set + merged order: exact
    assert torch.allclose(boxes[0, :n].cpu(), ref["proposal_boxes"], rtol=1e-5, atol=1e-3)
    assert float(boxes[0, :n, 0::2].max()) <= Hp and float(boxes[0, :n, 1::2].max()) <= Wp


@pytest.mark.parametrize("batch", [2, 4, 5, 6, 10, 16])
def test_rpn_proposals_batched_every_cluster_size(ops, batch):
    """The top-k and the NMS run as clusters of 8 / 7 / 5 / 4 / 2 / 1 CTAs per (level, image) depending on the batch
    (pick_nms_cluster); a batch is B independent images: every image must give what it gives alone (batch 1 = clusters
    of 8, pinned against the oracle above), bit for bit, including heavy ties in one of them."""
    sizes = [(64, 96), (32, 48), (16, 24), (8, 12), (4, 6)]
    per_image = [_rpn_inputs(100 + i, sizes, quant=0.5 if i == 1 else None)[2] for i in range(batch)]
    heads = [torch.cat([per_image[i][l] for i in range(batch)]).contiguous() for l in range(5)]
    boxes, scores, counts, dbg = ops.rpn_proposals(heads, clip_x=256.0, clip_y=384.0)
    torch.cuda.synchronize()
    for i in range(batch):
        b1, s1, c1, d1 = ops.rpn_proposals(per_image[i], clip_x=256.0, clip_y=384.0)
        torch.cuda.synchronize()
        n = int(c1[0])
        assert int(counts[i]) == n
        assert torch.equal(scores[i, :n], s1[0, :n]) and torch.equal(boxes[i, :n], b1[0, :n])
        assert torch.equal(dbg["cand_count"][i], d1["cand_count"][0])
        assert torch.equal(dbg["cand_scores"][i], d1["cand_scores"][0])
        assert torch.equal(dbg["cand_keep"][i], d1["cand_keep"][0])


def test_rpn_topk_with_many_ties(ops):
    sizes = [(40, 64), (20, 32), (10, 16), (5, 8), (3, 4)]
    logits, deltas, heads = _rpn_inputs(12, sizes, quant=0.5)           # heavy duplication of logit values
    lg, _ = O.rpn_flatten(logits, deltas)
    _, _, _, dbg = ops.rpn_proposals(heads, clip_x=160.0, clip_y=256.0)
    torch.cuda.synchronize()
    for l, s in enumerate(lg):
        k = min(1000, s.shape[1])
        ts, _ = s[0].topk(k)
        assert torch.equal(dbg["cand_scores"][0, l, :k].cpu(), ts)      # the multiset of selected scores is exact


# ----------------------------------------------------------------------------------------------- NMS
@pytest.mark.parametrize("n,thr,seed", [(1, 0.5, 0), (33, 0.5, 1), (1000, 0.7, 2), (1000, 0.5, 3), (777, 0.3, 4)])
def test_nms_keep_order_bitexact(ops, n, thr, seed):
    g = torch.Generator().manual_seed(seed)
    xy = torch.rand(n, 2, generator=g) * 300
    wh = torch.rand(n, 2, generator=g) * 120
    boxes = torch.cat([xy, xy + wh], 1)
    if n > 10:
        boxes[3] = boxes[2]                        # duplicates
        boxes[5, 2:] = boxes[5, :2]                # zero-area box: 0/0 = NaN never suppresses
        boxes[7] = boxes[5]
    scores = torch.rand(n, generator=g)
    order = torch.sort(scores, descending=True, stable=True)[1]
    sb = boxes[order]
    ref_keep = O.nms(sb, scores[order], thr)       # indices into the sorted list, ascending
    got = ops.nms_sorted(sb.cuda(), thr).cpu()
    assert torch.equal(torch.nonzero(got).squeeze(1), ref_keep)


def test_nms_threshold_boundary_and_degenerate_sets(ops):
    """The device kernel only computes the IEEE quotient near the threshold; these sets sit on it: integer-grid
    boxes whose IoU is exactly 1/2, 1/3, 2/3 (kept when == thr, torchvision uses a strict >), IoUs one ulp either
    side, 1000 identical boxes (one survivor), a suppression chain, and huge coordinates."""
    def run(boxes, thr):
        scores = torch.arange(boxes.shape[0], 0, -1).float()           # already sorted
        ref = O.nms(boxes, scores, thr)
        got = ops.nms_sorted(boxes.cuda(), thr).cpu()
        assert torch.equal(torch.nonzero(got).squeeze(1), ref), (thr, boxes.shape)

    # pairs (0,0,w,1) and (s,0,s+w,1): inter = w-s, union = w+s  ->  IoU = (w-s)/(w+s)
    rows = []
    for w, s_ in [(3, 1), (2, 1), (5, 1), (4, 2), (6, 2), (7, 3), (9, 3), (100, 50), (3, 0)]:
        y = 10.0 * len(rows)
        rows += [[0.0, y, float(w), y + 1.0], [float(s_), y, float(s_ + w), y + 1.0]]
    grid = torch.tensor(rows)
    for thr in (0.5, 1.0 / 3.0, 2.0 / 3.0, 0.7, 0.3, 0.0, 1.0):
        run(grid, thr)
    # thresholds one ulp either side of exactly representable IoUs
    for base in (0.5, 0.25, 0.75):
        t = torch.tensor(base)
        for thr in (float(torch.nextafter(t, torch.tensor(0.0))), float(torch.nextafter(t, torch.tensor(1.0)))):
            run(grid, thr)
    g = torch.Generator().manual_seed(5)
    # dense random set with many IoUs close to the threshold: jittered copies of a few boxes
    base = torch.tensor([[10.0, 10.0, 110.0, 60.0], [50.0, 20.0, 150.0, 90.0], [0.0, 0.0, 40.0, 200.0]])
    jit = base[torch.randint(0, 3, (1000,), generator=g)] + torch.rand(1000, 4, generator=g) * 60.0
    jit[:, 2:] = torch.maximum(jit[:, 2:], jit[:, :2])
    for thr in (0.5, 0.7):
        run(jit, thr)
    run(base[:1].repeat(1000, 1), 0.5)                                          # all identical
    chain = torch.stack([torch.tensor([float(i), 0.0, float(i) + 10.0, 10.0]) for i in range(1000)])
    run(chain, 0.7)                                                             # sliding chain of overlaps
    run(jit * 1.0e18, 0.5)                                                      # areas ~1e40: overflow to inf
    run(jit * 1.0e-22, 0.5)                                                     # areas in the denormal range


# ----------------------------------------------------------------------------------------------- ROIAlign
def test_roi_align_multilevel_bitexact(ops):
    g = torch.Generator().manual_seed(21)
    shapes = [(50, 84), (25, 42), (13, 21), (7, 11)]
    feats = [bf16(torch.randn(2, 256, h, w, generator=g)) for h, w in shapes]
    n = 64
    xy = torch.rand(n, 2, generator=g) * torch.tensor([300.0, 180.0])
    wh = torch.exp(torch.rand(n, 2, generator=g) * 5.0 + 1.0)           # ~3 .. 400 px: all four levels
    boxes = torch.cat([xy, xy + wh], 1)
    boxes[0] = torch.tensor([-40.0, -20.0, 30.0, 25.0])
    boxes[1] = torch.tensor([10.0, 10.0, 10.0, 10.0])
    boxes[2] = torch.tensor([320.0, 190.0, 900.0, 700.0])
    bidx = (torch.arange(n) % 2).float()
    rois = torch.cat([bidx[:, None], boxes], 1)
    levels = O.assign_boxes_to_levels(boxes)
    assert len(torch.unique(levels)) == 4
    scales = [1 / 4, 1 / 8, 1 / 16, 1 / 32]
    ref = torch.zeros(n, 256, 7, 7)
    for i in range(n):
        l = int(levels[i])
        b = int(bidx[i])
        r = torch.cat([torch.zeros(1, 1), boxes[i:i + 1]], 1)
        ref[i] = O.roi_align(feats[l][b:b + 1], r, 7, scales[l])[0]
    fe = [nhwc_bf16_cuda(f) for f in feats]
    got32 = ops.roi_align(fe, rois.cuda(), 7, scales, out_fp32=True)
    got16 = ops.roi_align(fe, rois.cuda(), 7, scales, out_fp32=False)
    torch.cuda.synchronize()
    g32 = got32.permute(0, 3, 1, 2).cpu()
    # identical sampling indices, weights, products and sum order (torchvision's CPU kernel has no FMA): bit-identical
    assert torch.equal(g32, ref), (float((g32 - ref).abs().max()), frac_equal(g32, ref))
    assert torch.equal(got16.permute(0, 3, 1, 2).float().cpu(), bf16(ref))


def test_roi_align_single_level_28(ops):
    g = torch.Generator().manual_seed(22)
    feat = bf16(torch.randn(1, 256, 60, 84, generator=g))
    n = 9
    xy = torch.rand(n, 2, generator=g) * torch.tensor([250.0, 150.0])
    wh = torch.rand(n, 2, generator=g) * 150 + 2
    rois = torch.cat([torch.zeros(n, 1), xy, xy + wh], 1)
    ref = O.roi_align(feat, rois, 28, 0.25)
    nv = torch.tensor([7], dtype=torch.int32, device="cuda")
    got = ops.roi_align([nhwc_bf16_cuda(feat)], rois.cuda(), 28, [0.25], out_fp32=True, n_rois=nv)
    torch.cuda.synchronize()
    g32 = got.permute(0, 3, 1, 2).cpu()
    assert torch.equal(g32[:7], ref[:7]), (float((g32[:7] - ref[:7]).abs().max()), frac_equal(g32[:7], ref[:7]))
    assert bool((g32[7:] == 0).all())              # rows past the device-side count are not touched


# ----------------------------------------------------------------------------------------------- box head tail
def test_box_predict_matches_oracle(ops):
    spec = O.SPECS["densepose_rcnn_R_50_FPN_s1x"]
    g = torch.Generator().manual_seed(31)
    R, n_prop = 1000, 940
    xy = torch.rand(R, 2, generator=g) * torch.tensor([1000.0, 600.0])
    wh = torch.rand(R, 2, generator=g) * 300 + 4
    props = torch.cat([xy, xy + wh], 1)
    cls = torch.randn(R, 2, generator=g) * 1.5
    dl = torch.randn(R, 4, generator=g) * 0.5
    head = torch.zeros(R, 16)
    head[:, :2], head[:, 2:6] = cls, dl
    boxes = O.apply_deltas(dl[:n_prop], props[:n_prop], (10.0, 10.0, 5.0, 5.0))
    probs = F.softmax(cls[:n_prop], dim=-1)
    det = O.fast_rcnn_inference_single_image(boxes, probs, torch.tensor([1344, 800]), spec)
    H0, W0, Hr, Wr = 480, 800, 800, 1333
    for k in ("pred_densepose_coarse_segm", "pred_densepose_fine_segm", "pred_densepose_u", "pred_densepose_v"):
        det[k] = torch.zeros(len(det["scores"]), 1)
    post = O.detector_postprocess(det, H0, W0, (0, 1344 - Wr, 0, 800 - Hr))
    raw, out_boxes, scores, count = ops.box_predict(
        head.cuda(), props[None].cuda().contiguous(), torch.tensor([n_prop], dtype=torch.int32, device="cuda"),
        spec.score_thresh, spec.nms_test, spec.dets_per_image, W0 / Wr, H0 / Hr, float(W0), float(H0))
    torch.cuda.synchronize()
    d = int(count[0])
    assert d == len(det["scores"])
    assert torch.allclose(scores[0, :d].cpu(), det["scores"], rtol=0, atol=2e-6)      # softmax: expf ulps
    assert torch.allclose(raw[0, :d].cpu(), det["pred_boxes"], rtol=1e-5, atol=2e-3)  # unclipped (quirk 2)
    assert torch.allclose(out_boxes[0, :d].cpu(), post["pred_boxes"], rtol=1e-5, atol=2e-3)
    assert float(out_boxes[0, :d, 0::2].max()) <= W0 and float(out_boxes[0, :d, 1::2].max()) <= H0


def test_box_predict_no_detections(ops):
    head = torch.zeros(1000, 16)
    head[:, 1] = 10.0                                                   # background wins everywhere
    props = torch.rand(1, 1000, 4) * 100
    props[..., 2:] += props[..., :2]
    raw, boxes, scores, count = ops.box_predict(head.cuda(), props.cuda().contiguous(),
                                                torch.tensor([1000], dtype=torch.int32, device="cuda"),
                                                0.3, 0.5, 100, 1.0, 1.0, 800.0, 600.0)
    torch.cuda.synchronize()
    assert int(count[0]) == 0


# ----------------------------------------------------------------------------------------------- DeepLab pieces
@pytest.mark.parametrize("C", [256, 512])
def test_groupnorm_relu(ops, C):
    g = torch.Generator().manual_seed(C)
    x = bf16(torch.randn(5, C, 28, 28, generator=g) * 2 + 0.3)
    gamma, beta = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g) * 0.2
    ref = bf16(F.relu(F.group_norm(x, 32, gamma, beta, eps=1e-5)))       # deeplab.py:45,70-73
    xin = x.permute(0, 2, 3, 1).reshape(5, 784, C).contiguous().to(torch.bfloat16).cuda()
    got = ops.groupnorm_relu(xin, gamma.cuda(), beta.cuda()).float().cpu().view(5, 28, 28, C).permute(0, 3, 1, 2)
    assert rel_l2(got, ref) < 3e-3 and frac_equal(got, ref) > 0.97


def test_groupnorm_relu_into_concat_slice_with_count(ops):
    """The ASPP branches write their normalised output straight into a channel slice of the 1280-channel concat
    (deeplab.py:141); ROIs past the device-side count stay untouched; two runs are bit-identical."""
    g = torch.Generator().manual_seed(9)
    R, C = 7, 256
    x = bf16(torch.randn(R, C, 28, 28, generator=g) * 1.5 - 0.2)
    gamma, beta = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g) * 0.2
    ref = bf16(F.relu(F.group_norm(x, 32, gamma, beta, eps=1e-5)))
    xin = x.permute(0, 2, 3, 1).reshape(R, 784, C).contiguous().to(torch.bfloat16).cuda()
    nv = torch.tensor([5], dtype=torch.int32, device="cuda")
    outs = []
    for _ in range(2):
        cat = torch.full((R, 784, 1280), -2.0, dtype=torch.bfloat16, device="cuda")
        ops.groupnorm_relu(xin, gamma.cuda(), beta.cuda(), out=cat[:, :, 512:768], n_valid=nv)
        torch.cuda.synchronize()
        outs.append(cat)
    assert torch.equal(outs[0], outs[1])
    cat = outs[0].float().cpu()
    got = cat[:5, :, 512:768].reshape(5, 28, 28, C).permute(0, 3, 1, 2)
    assert rel_l2(got, ref[:5]) < 3e-3 and frac_equal(got, ref[:5]) > 0.97
    assert bool((cat[5:] == -2.0).all()) and bool((cat[:, :, :512] == -2.0).all()) and bool((cat[:, :, 768:] == -2.0).all())


def test_avgpool_and_broadcast_gn(ops):
    g = torch.Generator().manual_seed(41)
    x = bf16(torch.randn(4, 256, 28, 28, generator=g))
    xin = x.permute(0, 2, 3, 1).reshape(4, 784, 256).contiguous().to(torch.bfloat16).cuda()
    pooled = ops.avgpool(xin)
    ref = bf16(F.adaptive_avg_pool2d(x, 1)[:, :, 0, 0])                   # deeplab.py:99
    assert rel_l2(pooled.float().cpu(), ref) < 3e-3
    gamma, beta = torch.rand(256, generator=g) + 0.5, torch.randn(256, generator=g) * 0.2
    got = ops.groupnorm_relu(pooled.view(4, 1, 256), gamma.cuda(), beta.cuda(), out_hw=784)
    refb = bf16(F.relu(F.group_norm(pooled.float().cpu().view(4, 256, 1, 1), 32, gamma, beta, eps=1e-5)))
    refb = refb.expand(4, 256, 28, 28)                                   # bilinear from 1x1 == broadcast (deeplab.py:109)
    assert rel_l2(got.float().cpu().view(4, 28, 28, 256).permute(0, 3, 1, 2), refb) < 3e-3


# ----------------------------------------------------------------------------------------------- predictor tail + extractor
@pytest.mark.parametrize("S,kc", [(56, 2), (28, 15), (28, 2)])
def test_predictor_upsample_is_bit_exact(ops, S, kc):
    """chart.py:72-74: fp32 bilinear x2 of the deconv output, bit-identical to ATen's CPU kernel — the separable one for
    the 112x112 outputs, the channels-last one (8-lane vector body + scalar tail per tensor) for the legacy 56x56."""
    g = torch.Generator().manual_seed(S)
    C = kc + 75
    cpad = (C + 15) // 16 * 16
    low = torch.randn(3, C, S, S, generator=g)
    parts = torch.split(low, [kc, 25, 25, 25], dim=1)                       # the reference upsamples four tensors
    ref = torch.cat([torch.from_numpy(AI.upsample_bilinear_f32(t.contiguous().numpy(), 2 * S, 2 * S, 2.0, 2.0)) for t in parts], 1)
    if X86:
        assert torch.equal(ref, torch.cat([F.interpolate(t, scale_factor=2.0, mode="bilinear", align_corners=False) for t in parts], 1))
    lin = torch.zeros(3, S, S, cpad)
    lin[..., :C] = low.permute(0, 2, 3, 1)
    lin = lin.view(3, S // 2, 2, S // 2, 2, cpad).permute(0, 2, 4, 5, 1, 3)    # [R, py, px, c, S/2, S/2]
    outs = ops.predictor_upsample(lin.cuda().contiguous(), (kc, 25, 25, 25))
    torch.cuda.synchronize()
    got = torch.cat([o.cpu() for o in outs], dim=1)
    assert torch.equal(got, ref), float((got - ref).abs().max())


@pytest.mark.parametrize("kc", [2, 15])
def test_dp_resample_is_bit_exact(ops, kc):
    """visualizer.py:10-56 on the device: labels (integer) and U/V must be IDENTICAL to the reference extractor,
    including ATen's kernel switch at h + w <= 128 and its vector / tail split (boxes 1, 4, 6-8 are small)."""
    g = torch.Generator().manual_seed(51 + kc)
    D, S = 9, 112
    low = [torch.randn(D, c, 14, 14, generator=g) for c in (kc, 25, 25, 25)]
    coarse, fine, u, v = [F.interpolate(t, size=(S, S), mode="bicubic", align_corners=False) for t in low]
    boxes = torch.tensor([[10.2, 20.7, 150.9, 300.1], [0.0, 0.0, 0.4, 0.3], [5.5, 5.5, 260.0, 90.2],
                          [100.0, 50.0, 131.9, 400.0], [7.0, 9.0, 119.0, 121.0], [3.3, 4.4, 60.6, 30.1],
                          [0.0, 0.0, 64.2, 64.9], [1.0, 1.0, 66.0, 65.0], [2.0, 2.0, 3.5, 130.0]])
    inst = {"pred_boxes": boxes, "pred_densepose_coarse_segm": coarse, "pred_densepose_fine_segm": fine,
            "pred_densepose_u": u, "pred_densepose_v": v}
    ref, ref_xywh = O.extract_results(inst, restated=True)                  # visualizer.py:46-56
    if X86:
        live, _ = O.extract_results(inst)                                   # F.interpolate itself
        for a, b in zip(ref, live):
            assert torch.equal(a["labels"], b["labels"]) and torch.equal(a["uv"], b["uv"])
    for u8 in (False, True):
        got, xywh = ops.dp_resample(coarse.cuda(), fine.cuda(), u.cuda(), v.cuda(), boxes.cuda(), labels_u8=u8)
        torch.cuda.synchronize()
        assert torch.equal(xywh.cpu(), ref_xywh)
        for r, q in zip(ref, got):
            assert r["labels"].shape == q["labels"].shape and q["labels"].dtype == (torch.uint8 if u8 else torch.int64)
            assert torch.equal(r["labels"], q["labels"].cpu().long())
            assert torch.equal(r["uv"], q["uv"].cpu())
