"""CPU-side checks: the C-ABI library loads and exports everything include/dpb200.h declares, the config
reader, the weight packer's re-layouts, the scriptable predictor, and frame sharding over gloo (world 2)."""
import io
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import REFERENCE, ROOT, have_reference


# ------------------------------------------------------------------------------------------- C ABI
def test_library_exports_every_declared_symbol():
    from densepose_torchscript_b200 import _lib
    header = open(os.path.join(ROOT, "include", "dpb200.h")).read()
    declared = set(re.findall(r"\b(dpb200_[a-z0-9_]+)\s*\(", header))
    declared -= {"dpb200_pack_conv_weight"}                 # mentioned in a comment only
    assert len(declared) >= 25
    for name in sorted(declared):
        assert hasattr(_lib.lib, name), f"{name} is declared in dpb200.h but not exported by libdpb200.so"
    assert declared == set(_lib.EXPORTS)
    assert _lib.lib.dpb200_abi_version() == 4
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    for name in declared:
        assert re.search(rf"\bT {name}\b", out), name


def test_no_device_means_loud_failure():
    from densepose_torchscript_b200 import _lib
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    assert _lib.lib.dpb200_device_ok() == 0
    with pytest.raises(_lib.DPB200Error):
        _lib.require_device()
    from densepose_torchscript_b200 import ops
    with pytest.raises(_lib.DPB200Error):
        ops.conv2d(torch.zeros(1, 4, 4, 64, dtype=torch.bfloat16), torch.zeros(16, 64, dtype=torch.bfloat16), None, 1, 1)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "densepose_torchscript_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), fn
    for fn in ("export.py", "run.py"):
        assert not re.search(r"^\s*(from|import)\s+oracle\b", open(os.path.join(ROOT, fn)).read(), re.M), fn


# ------------------------------------------------------------------------------------------- config
def test_builtin_specs_match_oracle_specs():
    from densepose_torchscript_b200.config import BUILTIN
    from oracle import densepose_oracle as O
    assert set(BUILTIN) == set(O.SPECS)
    for k, s in BUILTIN.items():
        o = O.SPECS[k]
        for f in ("depth", "head", "decoder_on", "pooler_res", "coarse_ch", "score_thresh", "nms_test", "dets_per_image",
                  "min_size", "max_size", "rpn_pre_topk", "rpn_post_topk", "rpn_nms", "pixel_mean", "pixel_std", "input_format",
                  "uv_confidence", "segm_confidence", "extra_heads"):
            assert getattr(s, f) == getattr(o, f), (k, f)


@pytest.mark.skipif(not have_reference(), reason="reference yaml files not present")
def test_yaml_reader_on_reference_configs():
    from densepose_torchscript_b200.config import BUILTIN, spec_from_yaml
    for name, spec in BUILTIN.items():
        y = spec_from_yaml(os.path.join(REFERENCE, "configs", name + ".yaml"))
        assert y == spec, name
    wc = spec_from_yaml(os.path.join(REFERENCE, "configs", "densepose_rcnn_R_50_FPN_WC1_s1x.yaml"), min_score=0.5, nms_thresh=0.4)
    assert (wc.head, wc.decoder_on, wc.pooler_res, wc.score_thresh, wc.nms_test) == ("v1convx", True, 28, 0.5, 0.4)


@pytest.mark.skipif(not have_reference(), reason="reference yaml files not present")
def test_yaml_reader_rejects_what_the_engine_does_not_compute():
    """Keys that change the numerics are validated, not ignored: CSE / HRNet / evolution yamls (other predictor,
    backbone, ROI heads or ROIAlignV2 poolers) raise instead of exporting a model that computes something else;
    WC* yamls are read with their confidence heads recorded."""
    from densepose_torchscript_b200.config import spec_from_yaml
    cfgs = os.path.join(REFERENCE, "configs")
    for bad, why in (("cse/densepose_rcnn_R_50_FPN_s1x.yaml", "PREDICTOR_NAME"),
                     ("HRNet/densepose_rcnn_HRFPN_HRNet_w32_s1x.yaml", "BACKBONE"),
                     ("evolution/densepose_R_50_FPN_DL_WC1M_3x_Atop10P_CA.yaml", "POOLER_TYPE")):
        with pytest.raises(ValueError, match=why):
            spec_from_yaml(os.path.join(cfgs, bad))
    with pytest.raises(FileNotFoundError, match="builtin names"):
        spec_from_yaml(os.path.join(cfgs, "densepose_rcnn_R_50_FPN_s1x_typo.yaml"))
    wc2m = spec_from_yaml(os.path.join(cfgs, "densepose_rcnn_R_50_FPN_WC2M_s1x.yaml"))
    assert (wc2m.uv_confidence, wc2m.segm_confidence) == ("indep_aniso", True)
    wc1 = spec_from_yaml(os.path.join(cfgs, "densepose_rcnn_R_101_FPN_DL_WC1_s1x.yaml"))
    assert (wc1.uv_confidence, wc1.segm_confidence, wc1.head, wc1.depth) == ("iid_iso", False, "deeplab", 101)


@pytest.mark.skipif(not have_reference(), reason="the reference visualizer is not present")
def test_visualizer_draws_exactly_like_the_reference():
    """End2EndVisualizer.draw == the reference's DensePoseResultsFineSegmentationVisualizer (visualizer.py:96-129) on
    the same extracted results: VIRIDIS, alpha blend per box, and the whole-frame fill of keep_bg=False (run.py:17)."""
    import importlib.util
    import cv2
    from densepose_torchscript_b200.extractor import End2EndVisualizer
    spec = importlib.util.spec_from_file_location("ref_visualizer", os.path.join(REFERENCE, "visualizer.py"))
    RV = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(RV)
    g = torch.Generator().manual_seed(0)
    img = torch.randint(0, 256, (200, 320, 3), dtype=torch.uint8, generator=g).numpy()
    boxes = torch.tensor([[10.2, 20.7, 150.9, 190.1], [0., 0., 0.4, 0.3], [5.5, 5.5, 260., 90.2], [100., 50., 131.9, 199.]])
    inst = {"pred_boxes": boxes, "pred_densepose_coarse_segm": torch.randn(4, 2, 112, 112, generator=g),
            "pred_densepose_fine_segm": torch.randn(4, 25, 112, 112, generator=g),
            "pred_densepose_u": torch.rand(4, 25, 112, 112, generator=g), "pred_densepose_v": torch.rand(4, 25, 112, 112, generator=g)}
    results, xywh = RV.DensePoseResultExtractor()(inst)
    for keep_bg in (True, False):
        want = RV.End2EndVisualizer(alpha=.7, keep_bg=keep_bg).visualize(img.copy(), inst)
        ours = End2EndVisualizer(alpha=.7, keep_bg=keep_bg)
        assert ours.cmap == cv2.COLORMAP_VIRIDIS
        assert np.array_equal(ours.draw(img.copy(), results, xywh), want)


# ------------------------------------------------------------------------------------------- weight packer
def _unpack(packed, kh, kw, cin, cout):
    w = packed[0].float().view(packed[3], kh, kw, packed[2])[:cout, :, :, :cin]
    return w.permute(0, 3, 1, 2)


def test_pack_folds_frozen_bn_and_accepts_aliases():
    from densepose_torchscript_b200 import synth
    from densepose_torchscript_b200.config import BUILTIN
    from densepose_torchscript_b200.weights import canonicalize, pack_state_dict
    spec = BUILTIN["densepose_rcnn_R_50_FPN_s1x"]
    sd = synth.make_state_dict(spec, 0)
    full = synth.add_aliases(sd, spec)
    only_alias = {k: v for k, v in full.items() if not (".res2." in k or "fpn_lateral" in k or "decoder.p3" in k or "body_conv_fcn" in k)}
    assert set(canonicalize(only_alias)) == set(sd)          # every duplicated registration maps back (quirk 10)
    packed = pack_state_dict(only_alias, spec, "cpu")
    p = "backbone.bottom_up.res3.1.conv2"
    scale = sd[p + ".norm.weight"] * (sd[p + ".norm.running_var"] + 1e-5).rsqrt()
    w_ref = (sd[p + ".weight"] * scale.view(-1, 1, 1, 1)).to(torch.bfloat16).float()
    b_ref = sd[p + ".norm.bias"] - sd[p + ".norm.running_mean"] * scale
    assert torch.equal(_unpack(packed[p], 3, 3, 128, 128), w_ref)
    assert torch.allclose(packed[p][1][:128], b_ref)
    # folded conv == conv + FrozenBN (batch_norm.py:54-62)
    x = torch.randn(1, 128, 9, 11)
    y_ref = F.batch_norm(F.conv2d(x, sd[p + ".weight"], padding=1), sd[p + ".norm.running_mean"], sd[p + ".norm.running_var"],
                         sd[p + ".norm.weight"], sd[p + ".norm.bias"], training=False, eps=1e-5)
    y = F.conv2d(x, sd[p + ".weight"] * scale.view(-1, 1, 1, 1), b_ref, padding=1)
    assert torch.allclose(y, y_ref, atol=1e-4)


def test_pack_special_layouts_are_exact_relayouts():
    from densepose_torchscript_b200 import synth
    from densepose_torchscript_b200.config import BUILTIN
    from densepose_torchscript_b200.weights import pack_state_dict
    spec = BUILTIN["densepose_rcnn_R_50_FPN_s1x"]
    sd = synth.make_state_dict(spec, 0)
    packed = pack_state_dict(sd, spec, "cpu")
    g = torch.Generator().manual_seed(0)
    # FC1: NHWC-flattened pooled features x permuted weight == reference flatten (c,y,x) x original weight
    pooled = torch.randn(3, 256, 7, 7, generator=g)
    w1 = packed["roi_heads.box_head.fc1"][0].float()
    ref = F.linear(pooled.flatten(1), sd["roi_heads.box_head.fc1.weight"].to(torch.bfloat16).float())
    got = F.linear(pooled.permute(0, 2, 3, 1).flatten(1), w1)
    assert torch.allclose(got, ref, atol=1e-3)
    # RPN head fusion: rows 0-2 objectness, 3-14 deltas, 15 zero
    rp = packed["proposal_generator.rpn_head.pred"]
    assert rp[3] == 16 and torch.equal(rp[0][15].float(), torch.zeros(256))
    assert torch.equal(rp[0][:3].float(), sd["proposal_generator.rpn_head.objectness_logits.weight"].view(3, 256).to(torch.bfloat16).float())
    # deconv phases: four 2x2 convs interleaved == ConvTranspose2d(k=4, s=2, p=1) (chart.py:45-59)
    x = torch.randn(2, 512, 6, 6, generator=g)
    names = ("ann_index_lowres", "index_uv_lowres", "u_lowres", "v_lowres")
    wt = torch.cat([sd[f"roi_heads.densepose_predictor.{n}.weight"] for n in names], 1).to(torch.bfloat16).float()
    bt = torch.cat([sd[f"roi_heads.densepose_predictor.{n}.bias"] for n in names], 0)
    ref = F.conv_transpose2d(x, wt, bt, stride=2, padding=1)
    out = torch.zeros_like(ref)
    pall = packed["roi_heads.densepose_predictor.phases"]          # the four phases stacked on Cout
    assert pall[0].shape == (320, 4 * 512) and pall[3] == 320 and pall[1].shape == (320,)
    for py in range(2):
        for px in range(2):
            i = py * 2 + px
            pk = (pall[0][80 * i:80 * i + 80], pall[1][80 * i:80 * i + 80], pall[2], 80)
            w = _unpack(pk, 2, 2, 512, 77)
            xp = F.pad(x, (1 if px == 0 else 0, 0 if px == 0 else 1, 1 if py == 0 else 0, 0 if py == 0 else 1))
            out[:, :, py::2, px::2] = F.conv2d(xp, w, pk[1][:77])
    assert torch.allclose(out, ref, atol=1e-3)
    # stem: 4x4 stride-1 conv over the 2x2 space-to-depth image (channels (dy, dx, c4)) == the 7x7 stride-2 pad-3 conv
    st = packed["backbone.bottom_up.stem.conv1"]
    assert st[0].shape == (64, 4 * 64)
    w4 = st[0].float().view(64, 4, 4, 16).permute(0, 3, 1, 2)                   # [co, (dy,dx,c4), ky', kx']
    w7 = sd["backbone.bottom_up.stem.conv1.weight"].float()
    bn = "backbone.bottom_up.stem.conv1.norm."
    scale = sd[bn + "weight"].float() * (sd[bn + "running_var"].float() + 1e-5).rsqrt()
    w7 = (w7 * scale.view(-1, 1, 1, 1)).to(torch.bfloat16).float()
    xs = torch.randn(1, 3, 20, 28, generator=g)
    ref7 = F.conv2d(xs, w7, stride=2, padding=3)
    x4 = torch.zeros(1, 4, 20, 28); x4[:, :3] = xs
    s2d = x4.view(1, 4, 10, 2, 14, 2).permute(0, 3, 5, 1, 2, 4).reshape(1, 16, 10, 14)    # channel = (dy*2+dx)*4 + c
    got4 = F.conv2d(F.pad(s2d, (2, 1, 2, 1)), w4)
    assert got4.shape == ref7.shape and torch.allclose(got4, ref7, atol=1e-4)


def test_deeplab_groupnorm_is_not_folded_into_the_conv():
    """DeepLab's `body_conv_fcnN.norm` is a GroupNorm (deeplab.py:45): its affine parameters stay a separate GN op, the
    packed conv is the raw bf16 weight without bias (only the backbone's FrozenBatchNorm2d layers are folded)."""
    from densepose_torchscript_b200 import synth
    from densepose_torchscript_b200.config import BUILTIN
    from densepose_torchscript_b200.weights import pack_state_dict
    spec = BUILTIN["densepose_rcnn_R_101_FPN_DL_s1x"]
    sd = synth.make_state_dict(spec, 0)
    packed = pack_state_dict(sd, spec, "cpu")
    k = "roi_heads.densepose_head.body_conv_fcn1"
    assert packed[k][1] is None
    w = sd[k + ".weight"].permute(0, 2, 3, 1).reshape(512, -1).to(torch.bfloat16)
    assert torch.equal(packed[k][0][:512, :w.shape[1]], w)
    assert torch.equal(packed[k + ".norm"][0], sd[k + ".norm.weight"].float())


def test_deeplab_rate56_is_its_centre_tap():
    x = torch.randn(2, 8, 28, 28)
    w = torch.randn(4, 8, 3, 3)
    assert torch.allclose(F.conv2d(x, w, padding=56, dilation=56), F.conv2d(x, w[:, :, 1:2, 1:2]), atol=1e-5)   # deeplab.py:35


# ------------------------------------------------------------------------------------------- predictor module
def test_predictor_scripts_saves_and_fails_loudly_on_cpu():
    from densepose_torchscript_b200 import synth
    from densepose_torchscript_b200.config import BUILTIN
    from densepose_torchscript_b200.predictor import DensePoseB200Predictor
    spec = BUILTIN["densepose_rcnn_R_50_FPN_s1x_legacy"]
    pred = DensePoseB200Predictor(spec, synth.add_aliases(synth.make_state_dict(spec, 0), spec)).eval()
    scripted = torch.jit.script(pred)
    assert "Tensor original_image, bool bgr=True) -> Dict(str, Tensor)" in str(scripted.forward.schema)
    buf = io.BytesIO()
    torch.jit.save(scripted, buf)
    buf.seek(0)
    loaded = torch.jit.load(buf).eval().half()
    assert (loaded.min_size, loaded.max_size, loaded.input_format) == (800, 1333, "BGR")
    assert loaded.dtype_probe.dtype == torch.float16 and loaded.weights.dtype == torch.uint8
    if not torch.cuda.is_available():
        with pytest.raises(Exception, match="no CPU implementation"):
            loaded(torch.zeros(8, 8, 3))


# ------------------------------------------------------------------------------------------- sharding
def test_shard_indices_cover_everything_once():
    from densepose_torchscript_b200.parallel import shard_indices
    for n in (0, 1, 7, 64, 65):
        for world in (1, 2, 8):
            for block in (1, 4, 8):
                seen = sorted(i for r in range(world) for i in shard_indices(n, r, world, block))
                assert seen == list(range(n))
    assert shard_indices(20, 1, 2, 4) == [4, 5, 6, 7, 12, 13, 14, 15]


_WORKER = r"""
import os, sys
sys.path.insert(0, sys.argv[1])
import torch.distributed as dist
from densepose_torchscript_b200.parallel import run_sharded
dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
items = list(range(37))
calls = []
def process(batch):
    calls.append(len(batch))
    return [{"frame": x, "rank": dist.get_rank(), "sq": x * x} for x in batch]
res = run_sharded(items, process, batch=4)
if dist.get_rank() == 0:
    assert [r["frame"] for r in res] == items and all(r["sq"] == r["frame"] ** 2 for r in res)
    assert {r["rank"] for r in res} == {0, 1}
    print("SHARD_OK", sum(1 for r in res if r["rank"] == 1))
else:
    assert res is None
assert max(calls) <= 4
dist.destroy_process_group()
"""


def test_run_sharded_gloo_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29531", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script), ROOT], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.PIPE, text=True) for r in range(2)]
    outs = [p.communicate(timeout=120) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "SHARD_OK 17" in outs[0][0]


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the driver's reference arm): one JSON line with impl/metric/value/unit, a
    cpu_baseline describing the run and an e2e object with zero copy bytes — on a small image so it runs in seconds."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--height", "96", "--width", "128"], capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "images/s" and line["value"] > 0
    assert line["higher_is_better"] is True and line["n_gpus"] == 1 and line["steps"] == 1
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_header_is_plain_c_and_matches_the_ctypes_structs(tmp_path):
    """include/dpb200.h compiles as C99 with gcc (no CUDA / C++ types cross the boundary) and every argument struct
    has the size (and last-field offset) of its ctypes mirror in _lib.py — an ABI drift between the header, the
    library and the binding would otherwise only show up as garbage on the GPU."""
    import ctypes as C
    from densepose_torchscript_b200 import _lib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pairs = [("dpb200_conv2d_args", _lib.Conv2dArgs), ("dpb200_preprocess_args", _lib.PreprocessArgs),
             ("dpb200_rpn_args", _lib.RpnArgs), ("dpb200_roi_align_args", _lib.RoiAlignArgs),
             ("dpb200_box_predict_args", _lib.BoxPredictArgs), ("dpb200_resample_args", _lib.ResampleArgs),
             ("dpb200_forward_io", _lib.ForwardIO), ("dpb200_model_config", _lib.ModelConfig),
             ("dpb200_weight", _lib.Weight)]
    src = tmp_path / "abi.c"
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "dpb200.h"', "int main(void) {"]
    for cname, ct in pairs:
        last = ct._fields_[-1][0]
        lines.append(f'  printf("{cname} %zu %zu\\n", sizeof({cname}), offsetof({cname}, {last}));')
    lines += ["  return 0;", "}"]
    src.write_text("\n".join(lines))
    exe = tmp_path / "abi"
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(root, "include"), str(src), "-o", str(exe)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = subprocess.run([str(exe)], capture_output=True, text=True).stdout.split("\n")
    got = {l.split()[0]: (int(l.split()[1]), int(l.split()[2])) for l in out if l.strip()}
    for cname, ct in pairs:
        last = ct._fields_[-1][0]
        assert got[cname] == (C.sizeof(ct), getattr(ct, last).offset), (cname, got[cname], C.sizeof(ct))


def test_caffe2_blob_names_convert_like_the_reference():
    """checkpoint.rename_caffe2_key == the reference's convert_c2_detectron_names on every blob of a DensePose R50-FPN
    (golden mapping generated from the reference by tests/golden/make_c2_names.py; re-derived live when the reference
    is present)."""
    import json
    from densepose_torchscript_b200.checkpoint import rename_caffe2_key
    here = os.path.dirname(os.path.abspath(__file__))
    golden = json.load(open(os.path.join(here, "golden", "c2_names.json")))["map"]
    assert len(golden) > 200
    for blob, name in golden.items():
        assert rename_caffe2_key(blob) == name, blob
    if have_reference():
        for p in (REFERENCE, os.path.join(ROOT, "oracle", "shims")):
            if p not in sys.path:
                sys.path.insert(0, p)
        from detectron2.checkpoint.c2_model_loading import convert_c2_detectron_names
        _, origin = convert_c2_detectron_names({b: torch.zeros(8, 2) for b in golden})
        assert {v: k for k, v in origin.items()} == golden


def test_caffe2_checkpoint_loads_into_the_reference_key_space(tmp_path):
    """A Detectron1-style {"blobs": ...} pickle (blob names, background row first in cls_score / bbox_pred, momentum
    blobs, AffineChannel scale/bias without running statistics) loads to exactly the tensors a detectron2-named
    state_dict holds, and packs to the same engine weights."""
    import json
    import pickle
    from densepose_torchscript_b200 import synth
    from densepose_torchscript_b200.checkpoint import load_checkpoint
    from densepose_torchscript_b200.config import BUILTIN
    from densepose_torchscript_b200.weights import pack_state_dict
    spec = BUILTIN["densepose_rcnn_R_50_FPN_s1x_legacy"]
    sd = synth.make_state_dict(spec, 0)
    for k in list(sd):                                   # Caffe2 models carry no running statistics
        if k.endswith("running_mean"):
            sd[k] = torch.zeros_like(sd[k])
        if k.endswith("running_var"):
            sd[k] = torch.ones_like(sd[k])
    here = os.path.dirname(os.path.abspath(__file__))
    golden = json.load(open(os.path.join(here, "golden", "c2_names.json")))["map"]
    blobs = {}
    for blob, short in golden.items():
        full = [k for k in sd if k == short or k.endswith("." + short)]
        assert len(full) == 1, (blob, short, full)
        v = sd[full[0]]
        if short.startswith("cls_score."):               # detectron2 keeps the background class last, Caffe2 first
            v = torch.cat([v[-1:], v[:-1]])
        if short.startswith("bbox_pred."):               # Caffe2 also regresses a (meaningless) background box
            v = torch.cat([torch.full_like(v[:4], 7.0), v])
        blobs[blob] = v.numpy()
        blobs[blob + "_momentum"] = np.zeros(1, dtype=np.float32)
    path = tmp_path / "DensePose_ResNet50_FPN_s1x-e2e.pkl"
    with open(path, "wb") as f:
        pickle.dump({"blobs": blobs, "cfg": "ignored"}, f)
    shapes = {k: tuple(v.shape) for k, v in sd.items()}
    got = load_checkpoint(str(path), shapes)
    stats = [k for k in sd if k.endswith("running_mean") or k.endswith("running_var")]
    assert set(got) == set(sd) - set(stats)
    for k in got:
        assert torch.equal(got[k], sd[k]), k
    a, b = pack_state_dict(got, spec, "cpu"), pack_state_dict(sd, spec, "cpu")
    assert set(a) == set(b)
    for k in a:
        assert torch.equal(a[k][0], b[k][0]), k
        assert (a[k][1] is None and b[k][1] is None) or torch.equal(a[k][1], b[k][1]), k


def test_bench_reference_arm_under_torchrun_only_rank0_works():
    """The driver launches the reference arm like the GPU arm (torchrun, N ranks): rank 0 alone measures and prints
    the one JSON line, the other ranks exit 0 without work."""
    import json
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29631", os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
           "--warmup", "0", "--height", "96", "--width", "128"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, lines
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["n_gpus"] == 2 and line["value"] > 0


def test_built_library_is_sm100a_tcgen05_code():
    """The shipped .so holds sm_100a SASS whose conv kernel uses the 5th-generation tensor-core path: tcgen05.mma
    (UTCHMMA, incl. the 2-CTA form), TMEM loads (LDTM), TMA tensor loads / stores (UTMALDG incl. im2col mode, UTMASTG),
    cluster-multicast commits (UTCBAR...MULTICAST) — and no legacy mma.sync (HMMA) anywhere."""
    import shutil
    from densepose_torchscript_b200 import _lib
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe):
        pytest.skip("cuobjdump not available")
    r = subprocess.run([exe, "-sass", _lib.LIB_PATH], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-500:]
    sass = r.stdout
    assert "sm_100a" in sass
    for needle in ("UTCHMMA", "UTCHMMA.2CTA", "LDTM", "UTMALDG.4D.IM2COL", "UTMALDG.2D.2CTA", "UTMASTG.2D", "MULTICAST"):
        assert needle in sass, needle
    assert "HMMA" not in sass.replace("UTCHMMA", "")
