"""The drop-in boundary on a real GPU: scripted predictor -> torch.ops.dpb200.forward -> C-ABI, and the
device-side result extractor, against the engine / oracle."""
import io

import pytest
import torch

from oracle import densepose_oracle as O
from oracle import weights as W

pytestmark = pytest.mark.gpu
NAME = "densepose_rcnn_R_50_FPN_s1x_legacy"


@pytest.fixture(scope="module")
def scripted():
    from densepose_torchscript_b200.config import BUILTIN
    from densepose_torchscript_b200.predictor import DensePoseB200Predictor
    sd = W.make_state_dict(O.SPECS[NAME], 0)
    pred = DensePoseB200Predictor(BUILTIN[NAME], W.add_aliases(sd, O.SPECS[NAME])).eval()
    buf = io.BytesIO()
    torch.jit.save(torch.jit.script(pred), buf)          # export.py:35-40
    buf.seek(0)
    return torch.jit.load(buf).eval().cuda(), sd         # run.py:18-26


def test_scripted_predictor_matches_engine_and_reference_contract(scripted):
    from densepose_torchscript_b200.config import BUILTIN
    from densepose_torchscript_b200.engine import Engine
    model, sd = scripted
    img = W.synthetic_image(200, 320, seed=11)
    out = model(img)                                       # CPU HWC float image, like run.py feeds it
    eng = Engine(BUILTIN[NAME], sd)
    ref = eng.forward_batch(img[None])[0]
    assert set(out) == {"image_size", "pred_boxes", "scores", "pred_classes", "pred_densepose_coarse_segm",
                        "pred_densepose_fine_segm", "pred_densepose_u", "pred_densepose_v"}
    for k in out:
        assert torch.equal(out[k], ref[k]), k
    d = len(out["scores"])
    assert out["pred_boxes"].dtype == torch.float32 and out["pred_classes"].dtype == torch.int64
    assert out["image_size"].tolist() == [200, 320] and out["image_size"].dtype == torch.int64
    assert out["pred_densepose_coarse_segm"].shape == (d, 15, 56, 56) and out["pred_densepose_u"].shape == (d, 25, 56, 56)
    assert all(v.is_contiguous() for v in out.values())
    out_chw = model(img.permute(2, 0, 1).contiguous())     # CHW input path (defaults.py:76-80)
    assert torch.equal(out_chw["pred_boxes"], out["pred_boxes"])
    with pytest.raises(Exception):
        model(torch.zeros(4, 32, 32))
    half = model.half()                                    # run.py:26
    oh = half(img)
    assert oh["scores"].dtype == torch.float16 and oh["pred_densepose_v"].dtype == torch.float16
    assert oh["pred_boxes"].dtype == torch.float32        # boxes stay fp32 (box_regression.py:84)
    model.float()


def test_extractor_matches_reference_visualizer_semantics(scripted):
    from densepose_torchscript_b200.extractor import DensePoseResultExtractor
    model, _ = scripted
    out = model(W.synthetic_image(200, 320, seed=12))
    results, xywh = DensePoseResultExtractor()(out)
    ref, ref_xywh = O.extract_results({k: v.float().cpu() for k, v in out.items()}, restated=True)     # visualizer.py:46-56
    assert torch.equal(xywh.cpu(), ref_xywh) and len(results) == len(ref) > 0
    for r, q in zip(ref, results):
        # integer part labels and the gathered U/V: identical to the CPU extractor (ATen's kernels bit for bit)
        assert torch.equal(r["labels"], q["labels"].cpu()) and torch.equal(r["uv"], q["uv"].cpu())


def test_run_py_video_batches_match_frame_by_frame(scripted, tmp_path):
    """run.py's video path: frames go through the engine in batches (HostPipeline + on-device extraction); the
    written video must be what the reference-style frame-by-frame loop over the exported module produces."""
    import importlib.util
    import os

    import cv2
    import numpy as np

    from densepose_torchscript_b200.extractor import End2EndVisualizer
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("dpb200_run", os.path.join(root, "run.py"))
    run = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(run)
    model, _ = scripted
    frames = [W.synthetic_image(160, 256, seed=40 + i).round().clamp(0, 255).to(torch.uint8).numpy() for i in range(7)]
    src = str(tmp_path / "clip.avi")
    w = cv2.VideoWriter(src, cv2.VideoWriter_fourcc(*"MJPG"), 25.0, (256, 160))
    for f in frames:
        w.write(f)
    w.release()
    cap = cv2.VideoCapture(src)
    decoded = []
    while True:
        ok, f = cap.read()
        if not ok:
            break
        decoded.append(f)
    cap.release()
    assert len(decoded) == 7
    vis = End2EndVisualizer(alpha=.7, inplace=False)
    want = [vis.visualize(f, model(torch.from_numpy(f))) for f in decoded]      # the reference's loop (run.py:42-57)

    got = []

    class Sink:                                    # stands in for cv2.VideoWriter: lossless capture of the frames
        def __init__(self, *a, **k):
            pass

        def write(self, frame):
            got.append(frame.copy())

        def release(self):
            pass

    real = cv2.VideoWriter
    cv2.VideoWriter = Sink
    try:
        cap = cv2.VideoCapture(src)
        n = run.run_video(cap, model, End2EndVisualizer(alpha=.7, inplace=True), str(tmp_path / "out.mp4"), 25.0, 3)
        cap.release()
    finally:
        cv2.VideoWriter = real
    assert n == 7 and len(got) == 7               # two full batches of 3 through the pipeline + a tail of 1
    for a, b in zip(got, want):
        assert a.shape == b.shape and np.array_equal(a, b)


def test_confidence_model_emits_its_extra_heads():
    """A WC* model: the scripted module returns the reference's eight keys plus one
    pred_densepose_<head> per confidence layer the config carries (chart_with_confidence.py:50-89), each the fp32
    ConvTranspose2d + bilinear x2 of the engine's own head output."""
    import torch.nn.functional as F

    from _util import nchw, rel_l2
    from densepose_torchscript_b200.config import BUILTIN
    from densepose_torchscript_b200.engine import Engine
    from densepose_torchscript_b200.predictor import DensePoseB200Predictor
    name = "densepose_rcnn_R_50_FPN_WC2M_s1x"
    spec = BUILTIN[name]
    assert [h for h, _ in spec.extra_heads] == ["sigma_2", "kappa_u", "kappa_v", "fine_segm_confidence", "coarse_segm_confidence"]
    sd = W.make_state_dict(O.SPECS[name], 0)
    pred = DensePoseB200Predictor(spec, W.add_aliases(sd, O.SPECS[name])).eval()
    buf = io.BytesIO()
    torch.jit.save(torch.jit.script(pred), buf)
    buf.seek(0)
    model = torch.jit.load(buf).eval().cuda()
    img = W.synthetic_image(200, 320, seed=11)
    out = model(img)
    d = len(out["scores"])
    assert d > 0
    for head, ch in spec.extra_heads:
        assert out["pred_densepose_" + head].shape == (d, ch, 112, 112)
    eng = Engine(spec, sd)
    ref = eng.forward_batch(img[None])[0]
    for k in out:
        assert torch.equal(out[k], ref[k]), k
    sess = eng.session(1, 200, 320, False)
    head_out = nchw(sess.tap("dp_head"))[:d]
    for head, _ in spec.extra_heads:
        p = "roi_heads.densepose_predictor." + head + "_lowres"
        low = F.conv_transpose2d(head_out, sd[p + ".weight"].to(torch.bfloat16).float(), sd[p + ".bias"], stride=2, padding=1)
        want = F.interpolate(low, scale_factor=2.0, mode="bilinear", align_corners=False)
        assert rel_l2(out["pred_densepose_" + head], want) < 2e-3, head
    # and through the host pipeline (count-aware D2H of every head)
    from densepose_torchscript_b200.engine import HostPipeline
    pipe = HostPipeline(eng, 1, 200, 320, False, depth=1)
    assert pipe.submit(img[None]) is None
    (got,) = pipe.drain()
    for k in ref:
        assert torch.equal(got[0][k], ref[k].cpu()), k
    pipe.close()


def test_export_and_run_clis_end_to_end(tmp_path):
    """The reference's two commands, unchanged (export.py:12-41, run.py:11-64): `export.py <cfg> <weights>` writes
    exported/<cfg>_fp32.pt, `run.py <model.pt> <image>` writes <image>_pred.png; the prediction drawn into the image is
    what the engine + the reference-identical visualiser produce for that uint8 image."""
    import os
    import subprocess
    import sys

    import cv2
    import numpy as np

    from densepose_torchscript_b200.config import BUILTIN
    from densepose_torchscript_b200.engine import Engine
    from densepose_torchscript_b200.extractor import End2EndVisualizer
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, PYTHONPATH=root + os.pathsep + os.environ.get("PYTHONPATH", ""))
    r = subprocess.run([sys.executable, os.path.join(root, "export.py"), NAME, ""], cwd=tmp_path, env=env,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    model = tmp_path / "exported" / (NAME + "_fp32.pt")
    assert model.exists()
    img = W.synthetic_image(200, 320, seed=21).round().clamp(0, 255).to(torch.uint8).numpy()
    src = tmp_path / "frame.png"
    cv2.imwrite(str(src), img)
    r = subprocess.run([sys.executable, os.path.join(root, "run.py"), str(model), str(src), "--fp32"], cwd=tmp_path, env=env,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    out = cv2.imread(str(tmp_path / "frame_pred.png"))
    assert out is not None and out.shape == img.shape
    eng = Engine(BUILTIN[NAME], W.make_state_dict(O.SPECS[NAME], 0))
    res = eng.forward_batch(torch.from_numpy(img)[None])[0]
    want = End2EndVisualizer(alpha=.7, keep_bg=False).visualize(img.copy(), res)
    assert np.array_equal(out, want)
    bad = subprocess.run([sys.executable, os.path.join(root, "export.py"), "densepose_rcnn_R_50_FPN_typo", ""], cwd=tmp_path,
                         env=env, capture_output=True, text=True, timeout=120)
    assert bad.returncode != 0 and "builtin" in (bad.stderr + bad.stdout)
