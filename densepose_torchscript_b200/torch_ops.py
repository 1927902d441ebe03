"""torch.ops.dpb200.* — the custom ops the scriptable predictor calls.

Registered from Python with torch.library (no torch headers in the build): the op bodies hand raw device
pointers and the current CUDA stream to libdpb200.so through ctypes.  Importing this module is what makes
`torch.jit.load` of an exported model work (run.py imports the package before loading, the same way the
reference's run.py imports torchvision "for its ops", run.py:7).

There is deliberately no CPU implementation: calling the op with CPU weights raises.
"""
import os
from collections import OrderedDict
from typing import Dict, List, Tuple

import torch

from . import _lib
from .config import ModelSpec
from .engine import Engine

_LIB = torch.library.Library("dpb200", "DEF")
_LIB.define("forward(Tensor image, bool bgr, Tensor blob, int[] table, str[] names, int[] cfg_i, float[] cfg_f, "
            "Tensor dtype_probe) -> Tensor[]")

_DT = {0: torch.bfloat16, 1: torch.float32}
# Engines behind exported modules, least recently used first. The key names the weight storage, the device AND the
# configuration lists, so two modules that share a blob but differ in thresholds get their own engine. The cache is
# bounded (DPB200_MAX_ENGINES, default 4): an engine pins its weights and up to `max_sessions` workspaces (0.9-7 GB
# each), so a process that keeps loading / moving modules must not accumulate them; `release_engines()` drops all.
_ENGINES: "OrderedDict[Tuple, Engine]" = OrderedDict()
_MAX_ENGINES = max(1, int(os.environ.get("DPB200_MAX_ENGINES", "4")))


def release_engines() -> None:
    """Frees every cached engine (weights stay with their modules; workspaces and sessions are released)."""
    _ENGINES.clear()


def spec_to_lists(spec: ModelSpec) -> Tuple[List[int], List[float]]:
    ci = [spec.depth, 0 if spec.head == "v1convx" else 1, int(spec.decoder_on), spec.pooler_res, spec.coarse_ch,
          spec.dets_per_image, spec.rpn_pre_topk, spec.rpn_post_topk, spec.min_size, spec.max_size,
          int(spec.input_format == "RGB"), {"": 0, "iid_iso": 1, "indep_aniso": 2}[spec.uv_confidence],
          int(spec.segm_confidence)]
    cf = [spec.score_thresh, spec.nms_test, spec.rpn_nms] + list(spec.pixel_mean) + list(spec.pixel_std)
    return ci, [float(x) for x in cf]


def lists_to_spec(ci: List[int], cf: List[float]) -> ModelSpec:
    return ModelSpec(name="exported", depth=ci[0], head="v1convx" if ci[1] == 0 else "deeplab", decoder_on=bool(ci[2]),
                     pooler_res=ci[3], coarse_ch=ci[4], dets_per_image=ci[5], rpn_pre_topk=ci[6], rpn_post_topk=ci[7],
                     min_size=ci[8], max_size=ci[9], input_format="RGB" if ci[10] else "BGR",
                     uv_confidence=("", "iid_iso", "indep_aniso")[ci[11]] if len(ci) > 11 else "",
                     segm_confidence=bool(ci[12]) if len(ci) > 12 else False,
                     score_thresh=cf[0], nms_test=cf[1], rpn_nms=cf[2], pixel_mean=tuple(cf[3:6]), pixel_std=tuple(cf[6:9]))


def pack_blob(packed) -> Tuple[torch.Tensor, List[int], List[str]]:
    """Flatten the packed parameter dict into one uint8 blob + an int table (8 ints per entry:
    off0, n0, dtype0, off1, n1, dtype1 (-1: absent), cin_pad, cout_pad)."""
    chunks, table, names, off = [], [], [], 0

    def add(t):
        nonlocal off
        raw = t.detach().contiguous().cpu().view(torch.uint8).reshape(-1)
        pad = (-off) % 256
        if pad:
            chunks.append(torch.zeros(pad, dtype=torch.uint8))
            off += pad
        chunks.append(raw)
        start = off
        off += raw.numel()
        return start, raw.numel()

    for name, (d0, d1, cin_pad, cout_pad) in packed.items():
        o0, n0 = add(d0)
        t0 = 0 if d0.dtype == torch.bfloat16 else 1
        if d1 is not None:
            o1, n1 = add(d1)
            t1 = 1
        else:
            o1, n1, t1 = 0, 0, -1
        table += [o0, n0, t0, o1, n1, t1, cin_pad, cout_pad]
        names.append(name)
    return torch.cat(chunks), table, names


def unpack_blob(blob: torch.Tensor, table: List[int], names: List[str]):
    packed = {}
    for i, name in enumerate(names):
        o0, n0, t0, o1, n1, t1, cin_pad, cout_pad = table[8 * i:8 * i + 8]
        d0 = blob[o0:o0 + n0].view(_DT[t0])
        d1 = blob[o1:o1 + n1].view(_DT[t1]) if t1 >= 0 else None
        packed[name] = (d0, d1, cin_pad, cout_pad)
    return packed


def _engine_for(blob: torch.Tensor, table: List[int], names: List[str], cfg_i: List[int], cfg_f: List[float]) -> Engine:
    key = (blob.data_ptr(), blob.get_device(), tuple(cfg_i), tuple(cfg_f))
    eng = _ENGINES.get(key)
    if eng is None:
        spec = lists_to_spec(cfg_i, cfg_f)
        packed = unpack_blob(blob, table, names)
        # conv weights are viewed as 2-D [cout_pad, K]
        packed = {k: (v[0].view(v[3], -1) if v[0].dtype == torch.bfloat16 else v[0], v[1], v[2], v[3]) for k, v in packed.items()}
        eng = Engine(spec, packed=packed, device=blob.device)
        eng._blob = blob    # keeps the storage alive for the engine's lifetime
        while len(_ENGINES) >= _MAX_ENGINES:
            _ENGINES.popitem(last=False)
    else:
        del _ENGINES[key]
    _ENGINES[key] = eng     # most recently used last
    return eng


def engine_of(module) -> Engine:
    """The Engine behind a (scripted or eager, already `.cuda()`) DensePoseB200Predictor: lets a caller that holds
    only the exported model batch frames through HostPipeline (run.py's video path)."""
    blob = module.weights
    if not blob.is_cuda:
        raise _lib.DPB200Error("engine_of: move the module to a B200 first (.cuda())")
    return _engine_for(blob, list(module.table), list(module.names), list(module.cfg_i), list(module.cfg_f))


def _forward_cuda(image, bgr, blob, table, names, cfg_i, cfg_f, dtype_probe):
    if not blob.is_cuda:
        raise _lib.DPB200Error("dpb200::forward has no CPU implementation: move the module to a B200 (.cuda())")
    eng = _engine_for(blob, table, names, cfg_i, cfg_f)
    if image.dim() != 3:
        raise ValueError("expected one image of shape (H, W, 3) or (3, H, W)")
    if image.shape[2] != 3:                                  # defaults.py:76-80
        if image.shape[0] != 3:
            raise AssertionError("Only 3 channels expected either in HWC or CHW format, got {}".format(tuple(image.shape)))
        image = image.permute(1, 2, 0)
    if image.dtype != torch.uint8:
        image = image.float()
    dt = dtype_probe.dtype if dtype_probe.dtype in (torch.float16, torch.bfloat16) else torch.float32
    # a `.half()` module gets its DensePose tensors as fp16 straight from the kernel (no conversion pass)
    res = eng.forward_batch(image.unsqueeze(0), bgr, out_half=(dt == torch.float16), copy=False)[0]   # cloned below
    out = [res["image_size"], res["pred_boxes"].clone(), res["scores"].to(dt, copy=True), res["pred_classes"],
           res["pred_densepose_coarse_segm"].to(dt, copy=True), res["pred_densepose_fine_segm"].to(dt, copy=True),
           res["pred_densepose_u"].to(dt, copy=True), res["pred_densepose_v"].to(dt, copy=True)]
    # WC* models: the confidence heads, in ModelSpec.extra_heads order (the scripted module names them)
    out += [res["pred_densepose_" + name].to(dt, copy=True) for name, _ in eng.spec.extra_heads]
    return out


# The image may arrive on the CPU (run.py feeds torch.from_numpy frames): dispatch on every backend and
# route by where the weights live.
_LIB.impl("forward", _forward_cuda, "CompositeExplicitAutograd")
