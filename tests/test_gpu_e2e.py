"""End-to-end parity of the engine (through the C-ABI session) against the oracle and against the golden
outputs of the real reference, plus batch / determinism properties at full size."""
import os

import pytest
import torch
import torch.nn.functional as F

from _util import bf16, coord_dist, match_detections, nchw, rel_l2
from conftest import GOLDEN
from oracle import densepose_oracle as O
from oracle import weights as W

pytestmark = pytest.mark.gpu

DP = ["pred_densepose_coarse_segm", "pred_densepose_fine_segm", "pred_densepose_u", "pred_densepose_v"]


def _engine(name):
    from densepose_torchscript_b200.config import BUILTIN
    from densepose_torchscript_b200.engine import Engine
    sd = W.make_state_dict(O.SPECS[name], 0)
    return Engine(BUILTIN[name], sd), sd


@pytest.fixture(scope="module")
def s1x():
    return _engine("densepose_rcnn_R_50_FPN_s1x")


def _compare_matched(res, ref, min_match, score_tol, box_tol, dp_tol, label_tol):
    ia, ib = match_detections(res["pred_boxes"], ref["pred_boxes"], box_tol)
    assert len(ia) >= min_match * max(len(ref["scores"]), 1), (len(ia), len(ref["scores"]))
    assert float((res["scores"][ia].cpu() - ref["scores"][ib]).abs().max()) < score_tol
    assert float((res["pred_boxes"][ia].cpu() - ref["pred_boxes"][ib]).abs().max()) < box_tol
    for k in DP:
        assert rel_l2(res[k][ia], ref[k][ib]) < dp_tol, k
    lab_e = res["pred_densepose_fine_segm"][ia].argmax(1).cpu()
    lab_r = ref["pred_densepose_fine_segm"][ib].argmax(1)
    agree = float((lab_e == lab_r).float().mean())
    assert agree > label_tol, agree
    return len(ia), agree


def test_stages_and_outputs_vs_bf16_oracle(s1x):
    """Tolerances (bf16 storage, fp32 accumulate, oracle rounds at the same points): backbone/FPN/decoder
    feature maps rel-L2 < 2e-2; matched detections: score < 1e-2, box < 1 px, DensePose tensors rel-L2 < 4e-2,
    fine-label pixel agreement > 97%."""
    eng, sd = s1x
    spec = O.SPECS["densepose_rcnn_R_50_FPN_s1x"]
    img = W.synthetic_image(240, 600, seed=3)
    taps = {}
    ref = O.forward(img, sd, spec, mode="bf16", taps=taps)
    sess = eng.session(1, 240, 600, False)
    sess.run(img[None].cuda().contiguous())
    torch.cuda.synchronize()
    assert (sess.hr, sess.wr, sess.hp, sess.wp) == (533, 1333, 544, 1344)
    for k in ("res2", "res3", "res4", "res5"):
        assert rel_l2(nchw(sess.tap(k)), taps["res"][k]) < 2e-2, k
    for k in ("p2", "p3", "p4", "p5"):
        assert rel_l2(nchw(sess.tap(k)), taps["feats"][k]) < 2e-2, k
    assert rel_l2(nchw(sess.tap("decoder")), taps["decoder"]) < 2e-2
    for l in range(5):
        h = sess.tap(f"rpn_head{l}")[0].float().cpu()
        assert rel_l2(h[..., :3].reshape(-1), taps["rpn_logits"][l][0]) < 3e-2
        assert rel_l2(h[..., 3:15].reshape(-1, 4), taps["rpn_deltas"][l][0]) < 3e-2
    # proposals: same set up to near-tie reordering of the top-k / NMS -> nearest-corner matching
    n = int(sess.tap("proposal_count").view(-1)[0])
    pb = sess.tap("proposal_boxes")[0, :n, :, 0].float().cpu()
    dist = coord_dist(pb, taps["proposals"]["proposal_boxes"])
    frac = float((dist.min(dim=1).values < 1.0).float().mean())
    assert frac > 0.7, frac
    res = sess.results()[0]
    assert abs(len(res["scores"]) - len(ref["scores"])) <= 3
    assert torch.equal(res["image_size"].cpu(), ref["image_size"])
    assert bool((res["scores"][:-1] >= res["scores"][1:]).all())
    _compare_matched(res, ref, min_match=0.85, score_tol=1e-2, box_tol=1.0, dp_tol=4e-2, label_tol=0.97)
    # deconv phases + fused predictor tail against conv_transpose2d on the engine's own head output
    d = len(res["scores"])
    head = nchw(sess.tap("dp_head"))[:d]
    low = F.conv_transpose2d(head, bf16(sd["roi_heads.densepose_predictor.u_lowres.weight"]),
                             sd["roi_heads.densepose_predictor.u_lowres.bias"], stride=2, padding=1)   # chart.py:55-59
    u_ref = F.interpolate(low, scale_factor=2.0, mode="bilinear", align_corners=False)
    assert rel_l2(res["pred_densepose_u"], u_ref) < 2e-3


def test_batch_images_are_independent_and_deterministic(s1x):
    """Batch is an engine extension (the reference is batch-1, rcnn.py:161): semantics = B independent calls."""
    eng, _ = s1x
    a, b = W.synthetic_image(240, 600, seed=3), W.synthetic_image(240, 600, seed=4)
    # results are views of the session's output buffers: clone before the session is reused
    single = [{k: v.clone() for k, v in eng.forward_batch(x[None])[0].items()} for x in (a, b)]
    batch = [{k: v.clone() for k, v in r.items()} for r in eng.forward_batch(torch.stack([a, b, a]))]
    for r, s in ((batch[0], single[0]), (batch[1], single[1]), (batch[2], single[0])):
        for k in s:
            assert torch.equal(r[k], s[k]), k
    again = eng.forward_batch(torch.stack([a, b, a]))
    for r, s in zip(again, batch):
        for k in s:
            assert torch.equal(r[k], s[k]), k


def test_graph_replay_equals_plain_launches(s1x):
    """The CUDA-graph replay (programmatic-dependent-launch edges between the convs captured into the graph) and plain
    stream launches of the same plan give bit-identical outputs, run after run."""
    from densepose_torchscript_b200.config import BUILTIN
    from densepose_torchscript_b200.engine import Engine
    eng, sd = s1x
    assert eng.use_graph
    plain = Engine(BUILTIN["densepose_rcnn_R_50_FPN_s1x"], sd, use_graph=False)
    imgs = torch.stack([W.synthetic_image(240, 600, seed=3), W.synthetic_image(240, 600, seed=9)])
    ref = [{k: v.clone() for k, v in r.items()} for r in plain.forward_batch(imgs)]
    for _ in range(3):                      # first call captures, the next ones replay
        got = eng.forward_batch(imgs)
        torch.cuda.synchronize()
        for r, s in zip(got, ref):
            for k in s:
                assert torch.equal(r[k], s[k]), k


@pytest.mark.parametrize("name,batch,size", [("densepose_rcnn_R_50_FPN_s1x", 1, (480, 640)), ("densepose_rcnn_R_50_FPN_s1x", 3, (240, 600)),
                                             ("densepose_rcnn_R_101_FPN_DL_s1x", 2, (240, 600)),
                                             ("densepose_rcnn_R_50_FPN_s1x_legacy", 2, (240, 600))])
def test_two_stream_schedule_equals_the_serial_one(name, batch, size, monkeypatch):
    """The session runs its launches as a two-branch graph (proposal / box chain and the small FPN / RPN levels on a side
    stream, explicit event edges, engine.cu run_ops). DPB200_SERIAL_SCHEDULE=1 runs the same launches in list order on one
    stream: every output and the intermediate maps must be bit-identical, run after run, with plain launches and with the
    captured graph - a missing edge is a race and shows up here as a difference."""
    from densepose_torchscript_b200.config import BUILTIN
    from densepose_torchscript_b200.engine import Engine
    spec = BUILTIN[name]
    sd = W.make_state_dict(O.SPECS[name], 0)
    imgs = torch.stack([W.synthetic_image(size[0], size[1], seed=40 + i) for i in range(batch)])
    monkeypatch.setenv("DPB200_SERIAL_SCHEDULE", "1")
    serial = Engine(spec, sd, use_graph=False)
    ref = [{k: v.clone() for k, v in r.items()} for r in serial.forward_batch(imgs)]
    sess = serial.session(batch, size[0], size[1], False)
    taps = {t: sess.tap(t).clone() for t in ("p2", "p3", "p4", "p5", "rpn_head0", "rpn_head3", "rpn_head4", "proposal_boxes")}
    monkeypatch.delenv("DPB200_SERIAL_SCHEDULE")
    for use_graph in (False, True):
        eng = Engine(spec, sd, use_graph=use_graph)
        for it in range(12):
            got = eng.forward_batch(imgs)
            torch.cuda.synchronize()
            for r, s in zip(got, ref):
                assert set(r) == set(s)
                for k in s:
                    assert torch.equal(r[k], s[k]), (use_graph, it, k)
        es = eng.session(batch, size[0], size[1], False)
        for t, v in taps.items():
            assert torch.equal(es.tap(t), v), (use_graph, t)


def test_full_size_properties(s1x):
    """BASELINE-size input (800x1333): shapes, ordering, clipping and the zero-detection path."""
    eng, _ = s1x
    img = W.synthetic_image(800, 1333, seed=1)
    res = eng.forward_batch(img[None])[0]
    d = len(res["scores"])
    assert 0 < d <= 100
    assert res["pred_boxes"].shape == (d, 4) and res["pred_densepose_fine_segm"].shape == (d, 25, 112, 112)
    assert res["pred_densepose_coarse_segm"].shape == (d, 2, 112, 112)
    assert bool((res["scores"] > 0.3).all()) and bool((res["scores"][:-1] >= res["scores"][1:]).all())
    bx = res["pred_boxes"]
    assert float(bx.min()) >= 0 and float(bx[:, 0::2].max()) <= 1333 and float(bx[:, 1::2].max()) <= 800
    assert bool(torch.isfinite(res["pred_densepose_u"]).all())
    # an all-zero image after mean subtraction still runs; a constant image typically yields few/no detections
    flat = torch.full((800, 1333, 3), 110.0)
    r0 = eng.forward_batch(flat[None])[0]
    assert r0["pred_boxes"].shape[1] == 4 and r0["pred_densepose_u"].shape[1:] == (25, 112, 112)
    assert len(r0["scores"]) == r0["pred_densepose_u"].shape[0]


def test_portrait_image_swapped_clip_quirk_on_the_device(s1x):
    """A portrait image (300x200 -> 1200x800 -> padded 1216x800): quirk 1 clips proposal x to [0, H_pad] and y to
    [0, W_pad] (rpn.py:339 vs structures.py:107-112); on a portrait input that lets boxes run past the image on the
    right. Engine vs the bf16 oracle on the same image: geometry, the clip extents, and matched detections."""
    eng, sd = s1x
    spec = O.SPECS["densepose_rcnn_R_50_FPN_s1x"]
    img = W.synthetic_image(300, 200, seed=5)
    taps = {}
    ref = O.forward(img, sd, spec, mode="bf16", taps=taps)
    sess = eng.session(1, 300, 200, False)
    sess.run(img[None].cuda().contiguous())
    torch.cuda.synchronize()
    assert (sess.hr, sess.wr, sess.hp, sess.wp) == (1200, 800, 1216, 800)
    n = int(sess.tap("proposal_count").view(-1)[0])
    pb = sess.tap("proposal_boxes")[0, :n, :, 0].float().cpu()
    ref_pb = taps["proposals"]["proposal_boxes"]
    assert float(pb[:, 0::2].max()) <= 1216.0 and float(pb[:, 1::2].max()) <= 800.0      # x <= H_pad, y <= W_pad
    assert float(pb[:, 0::2].max()) > 800.0 or float(ref_pb[:, 0::2].max()) <= 800.0     # the quirk bites when it does in the oracle
    assert abs(n - len(ref_pb)) <= 30
    res = sess.results()[0]
    assert torch.equal(res["image_size"].cpu(), ref["image_size"])
    ia, ib = match_detections(res["pred_boxes"], ref["pred_boxes"], 2.0)
    assert len(ia) >= 0.85 * len(ref["scores"])
    assert float((res["scores"][ia].cpu() - ref["scores"][ib]).abs().max()) < 1e-2


@pytest.mark.parametrize("h0,w0", [(64, 64), (33, 257), (257, 33), (101, 37)])
def test_ragged_image_sizes_run_like_the_oracle(s1x, h0, w0):
    """Extreme aspect ratios and tiny inputs (the reference takes any H x W): the resize geometry, the pyramid sizes
    (p6 = ceil(p5 / 2) down to a single row / column) and the result shapes follow the oracle; detections match."""
    eng, sd = s1x
    spec = O.SPECS["densepose_rcnn_R_50_FPN_s1x"]
    img = W.synthetic_image(h0, w0, seed=h0 + w0)
    ref = O.forward(img, sd, spec, mode="bf16")
    res = eng.forward_batch(img[None])[0]
    torch.cuda.synchronize()
    assert torch.equal(res["image_size"].cpu(), ref["image_size"]) and res["image_size"].tolist() == [h0, w0]
    assert abs(len(res["scores"]) - len(ref["scores"])) <= max(3, len(ref["scores"]) // 10)
    assert res["pred_densepose_u"].shape[1:] == (25, 112, 112)
    if len(ref["scores"]):
        ia, ib = match_detections(res["pred_boxes"], ref["pred_boxes"], 2.0 * max(h0, w0) / 200 + 0.5)
        assert len(ia) >= 0.7 * len(ref["scores"])
        bx = res["pred_boxes"]
        assert float(bx.min()) >= 0 and float(bx[:, 0::2].max()) <= w0 and float(bx[:, 1::2].max()) <= h0


def test_uint8_input_session(s1x):
    """uint8 frames (run.py:33-36) take ATen's fixed-point resize inside the session: the stem input equals the
    per-op kernel's (bit-exact against ATen in test_gpu_ops) and differs from the float path's, whose resize rounds
    differently; detections stay close. (Parity against the real reference on uint8 input: test_gpu_parity.)"""
    from densepose_torchscript_b200 import ops
    eng, _ = s1x
    spec = O.SPECS["densepose_rcnn_R_50_FPN_s1x"]
    img = W.synthetic_image(240, 600, seed=3).round().clamp(0, 255)
    u8 = img.to(torch.uint8)[None].cuda().contiguous()
    r8 = eng.forward_batch(u8)[0]
    s8 = eng.session(1, 240, 600, True)
    k = O.resize_scale(240, 600, spec)
    want, _ = ops.preprocess(u8, k, spec.pixel_mean, spec.pixel_std)
    torch.cuda.synchronize()
    assert torch.equal(s8.tap("stem_in"), want)
    rf = eng.forward_batch(img[None])[0]
    sf = eng.session(1, 240, 600, False)
    assert not torch.equal(sf.tap("stem_in"), want)
    ia, ib = match_detections(r8["pred_boxes"], rf["pred_boxes"], 4.0)
    assert len(ia) >= 0.6 * len(rf["scores"])


def test_half_outputs_are_the_rounded_fp32_outputs(s1x):
    """out_half sessions (the reference's `.half()` output contract, run.py:20-29): the kernel writes fp16 itself;
    every value must be the round-to-nearest-even fp16 of the fp32-output run, boxes / scores unchanged."""
    eng, _ = s1x
    img = W.synthetic_image(240, 600, seed=3)
    r32 = {k: v.clone() for k, v in eng.forward_batch(img[None])[0].items()}
    r16 = eng.forward_batch(img[None], out_half=True)[0]
    assert len(r32["scores"]) > 0
    assert torch.equal(r16["pred_boxes"], r32["pred_boxes"]) and torch.equal(r16["scores"], r32["scores"])
    for k in DP:
        assert r16[k].dtype == torch.float16
        assert torch.equal(r16[k], r32[k].half()), k


def test_pipelined_extractor_matches_the_extractor_on_full_outputs(s1x):
    """HostPipeline(extract=True) (run.py flow: forward -> per-box resample / argmax / UV gather on the device, only
    labels + uv cross PCIe) against DensePoseResultExtractor applied to the full output tensors: bit-equal."""
    from densepose_torchscript_b200.engine import HostPipeline
    from densepose_torchscript_b200.extractor import DensePoseResultExtractor
    eng, _ = s1x
    imgs = torch.stack([W.synthetic_image(240, 600, seed=3), W.synthetic_image(240, 600, seed=4)])
    full = [{k: v.clone() for k, v in r.items()} for r in eng.forward_batch(imgs)]
    for u8 in (True, False):
        pipe = HostPipeline(eng, 2, 240, 600, False, depth=2, extract=True, labels_u8=u8)
        assert pipe.submit(imgs) is None
        (res,) = pipe.drain()
        assert pipe.extract_d2h_bytes > 0
        for b in range(2):
            ref, xywh = DensePoseResultExtractor()(full[b])
            assert torch.equal(res[b]["pred_boxes"], full[b]["pred_boxes"].cpu())
            assert torch.equal(res[b]["boxes_xywh"], xywh.cpu())
            assert len(res[b]["densepose"]) == len(ref) > 0
            for got, want in zip(res[b]["densepose"], ref):
                assert got["labels"].dtype == (torch.uint8 if u8 else torch.int64)
                assert torch.equal(got["labels"].long(), want["labels"].cpu())
                assert torch.equal(got["uv"], want["uv"].cpu())
        pipe.close()


def test_pipeline_results_survive_the_next_submit_and_copies_are_count_aware(s1x):
    """HostPipeline with DISTINCT images per batch and a consumer that reads results late: what submit() returns stays
    valid until the NEXT submit() has returned (depth + 1 rotating pinned result sets; round 1 overwrote them from the
    same call). The DensePose rows come back count-aware: only the rows that hold detections cross PCIe."""
    from densepose_torchscript_b200.engine import HostPipeline
    eng, _ = s1x
    batches = [torch.stack([W.synthetic_image(240, 600, seed=20 + 2 * i), W.synthetic_image(240, 600, seed=21 + 2 * i)])
               for i in range(6)]
    want = [[{k: v.cpu() for k, v in r.items()} for r in eng.forward_batch(b)] for b in batches]
    pipe = HostPipeline(eng, 2, 240, 600, False, depth=2)
    got = []
    held = None
    returned = []
    for b in batches:
        r = pipe.submit(b)                       # issues the next copies and enqueues the next batch
        returned.append(r is not None)
        if held is not None:
            torch.cuda.synchronize()             # the new batch has certainly run: `held` must still be intact
            got.append([{k: v.clone() for k, v in d.items()} for d in held])
        held = r
        if r is not None:
            assert pipe.last_d2h_bytes == pipe.small_d2h_bytes + sum(len(d["scores"]) for d in r) * pipe.row_bytes
            assert pipe.last_d2h_bytes <= pipe.d2h_bytes
    if held is not None:                         # drain() reuses both result sets: take the last submit's results first
        got.append([{k: v.clone() for k, v in d.items()} for d in held])
    tail = pipe.drain()
    got += [[{k: v.clone() for k, v in d.items()} for d in r] for r in tail]
    assert returned == [False, False, False, True, True, True]      # results come back depth + 1 calls later
    assert len(got) == 6 and len(tail) == 3
    for g, w in zip(got, want):
        for a, b in zip(g, w):
            assert len(a["scores"]) == len(b["scores"]) > 0
            for k in b:
                assert torch.equal(a[k], b[k]), k
    pipe.close()


def test_1080p_frames_r101(tmp_path):
    """configs[3]: R_101_FPN_s1x on 1080x1920 video frames (-> 749x1333 -> padded 768x1344):
    geometry, shapes, clipping to the frame, uint8 frames and batch independence at that size."""
    eng, _ = _engine("densepose_rcnn_R_101_FPN_s1x")
    frames = torch.stack([W.synthetic_image(1080, 1920, seed=11), W.synthetic_image(1080, 1920, seed=12)])
    sess = eng.session(2, 1080, 1920, False)
    assert (sess.hr, sess.wr, sess.hp, sess.wp) == (749, 1333, 768, 1344)
    res = [{k: v.clone() for k, v in r.items()} for r in eng.forward_batch(frames)]
    for r in res:
        d = len(r["scores"])
        assert 0 < d <= 100 and r["pred_densepose_u"].shape == (d, 25, 112, 112)
        assert torch.equal(r["image_size"].cpu(), torch.tensor([1080, 1920]))
        bx = r["pred_boxes"]
        assert float(bx.min()) >= 0 and float(bx[:, 0::2].max()) <= 1920 and float(bx[:, 1::2].max()) <= 1080
        assert bool(torch.isfinite(r["pred_densepose_fine_segm"]).all())
    single = eng.forward_batch(frames[1:2])[0]
    for k in single:
        assert torch.equal(single[k], res[1][k]), k
    r8 = eng.forward_batch(frames[:1].round().clamp(0, 255).to(torch.uint8))[0]
    assert r8["pred_densepose_u"].shape[1:] == (25, 112, 112)


def test_session_cache_is_bounded(s1x):
    """A stream of differently sized images must not accumulate workspaces: the engine keeps max_sessions plain
    sessions and drops the least recently used."""
    from densepose_torchscript_b200.config import BUILTIN
    from densepose_torchscript_b200.engine import Engine
    eng = Engine(BUILTIN["densepose_rcnn_R_50_FPN_s1x"], packed=s1x[0].packed, max_sessions=2)
    for h, w in ((64, 96), (96, 64), (80, 80), (64, 96)):
        r = eng.forward_batch(W.synthetic_image(h, w, seed=h)[None])[0]
        assert r["image_size"].tolist() == [h, w]
        assert len(eng._sessions) <= 2
    assert (1, 64, 96, False, 0, False) in eng._sessions and (1, 80, 80, False, 0, False) in eng._sessions


@pytest.mark.parametrize("name", ["densepose_rcnn_R_50_FPN_s1x", "densepose_rcnn_R_101_FPN_DL_s1x"])
def test_no_detections_gives_correctly_shaped_empty_outputs(name):
    """D = 0: with a score threshold nothing passes, every ROI-side kernel sees a device-side count of
    zero (skipped tiles in the CTA-pair convs, the one-launch deconv, GroupNorm, the predictor tail) and the result
    is the reference's empty dict; a following normal run on the same engine is unaffected."""
    from dataclasses import replace

    from densepose_torchscript_b200.config import BUILTIN
    from densepose_torchscript_b200.engine import Engine
    sd = W.make_state_dict(O.SPECS[name], 0)
    spec = replace(BUILTIN[name], min_size=256, max_size=448)
    imgs = torch.stack([W.synthetic_image(128, 192, seed=5), W.synthetic_image(128, 192, seed=6)])
    none = Engine(replace(spec, score_thresh=0.99999), sd).forward_batch(imgs)
    torch.cuda.synchronize()
    for r in none:
        assert r["pred_boxes"].shape == (0, 4) and r["scores"].shape == (0,) and r["pred_classes"].shape == (0,)
        assert r["pred_densepose_coarse_segm"].shape == (0, spec.coarse_ch, 112, 112)
        for k in DP[1:]:
            assert r[k].shape == (0, 25, 112, 112)
        assert r["image_size"].tolist() == [128, 192]
    some = Engine(spec, sd).forward_batch(imgs)
    torch.cuda.synchronize()
    assert all(len(r["scores"]) > 0 for r in some)
    assert all(bool(torch.isfinite(r["pred_densepose_u"]).all()) for r in some)


def test_channel_order_quirk(s1x):
    """defaults.py:82-83: only an INPUT.FORMAT == "RGB" model flips the channels of a bgr=True input;
    with the default BGR config the `bgr` flag changes nothing."""
    from dataclasses import replace

    from densepose_torchscript_b200.config import BUILTIN
    from densepose_torchscript_b200.engine import Engine
    _, sd = s1x
    base = replace(BUILTIN["densepose_rcnn_R_50_FPN_s1x"], min_size=256, max_size=448)
    img = W.synthetic_image(128, 192, seed=12)[None]
    clone = lambda rs: [{k: v.clone() for k, v in r.items()} for r in rs]      # noqa: E731
    bgr_model = Engine(base, sd)
    a = clone(bgr_model.forward_batch(img, bgr=True))
    b = clone(bgr_model.forward_batch(img, bgr=False))
    for k in a[0]:
        assert torch.equal(a[0][k], b[0][k]), k
    rgb_model = Engine(replace(base, input_format="RGB"), sd)
    c = clone(rgb_model.forward_batch(img, bgr=True))                          # flipped inside
    d = clone(rgb_model.forward_batch(img.flip(-1).contiguous(), bgr=False))  # flipped by the caller
    for k in c[0]:
        assert torch.equal(c[0][k], d[0][k]), k
    e = clone(rgb_model.forward_batch(img, bgr=False))                         # RGB model, RGB input: no flip
    assert len(e[0]["scores"]) != len(c[0]["scores"]) or not torch.equal(e[0]["pred_boxes"], c[0]["pred_boxes"])
