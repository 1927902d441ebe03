"""Drop-in for the reference's `DefaultPredictor` (detectron2/engine/defaults.py:50-97).

Same call contract — `forward(original_image, bgr=True) -> Dict[str, Tensor]` with one image of shape
(H, W, 3) or (3, H, W) — and the module is `torch.jit.script`-able / `torch.jit.save`-able, so export.py and
run.py keep working unchanged in spirit.  The body is one custom op, `torch.ops.dpb200.forward`, which runs
the whole forward pass as hand-written sm_100a kernels (libdpb200.so).  Weights travel inside the module as
one packed uint8 buffer, so `.cuda()`, `.half()`, `.float()` (run.py:20-29) are safe: `.half()` only flips
the output dtype of scores / DensePose tensors, exactly like the reference.
"""
from typing import Dict, List

import torch
from torch import nn

from . import torch_ops  # noqa: F401  (registers torch.ops.dpb200)
from .config import ModelSpec
from .weights import pack_state_dict


class DensePoseB200Predictor(nn.Module):
    extra_names: List[str]          # (class-level annotation: TorchScript cannot infer the type of an empty list)

    def __init__(self, spec: ModelSpec, state_dict: Dict[str, torch.Tensor]):
        super().__init__()
        packed = pack_state_dict(state_dict, spec, "cpu")
        blob, table, names = torch_ops.pack_blob(packed)
        self.register_buffer("weights", blob)
        self.register_buffer("dtype_probe", torch.zeros(1, dtype=torch.float32))
        self.table: List[int] = table
        self.names: List[str] = names
        ci, cf = torch_ops.spec_to_lists(spec)
        self.cfg_i: List[int] = ci
        self.cfg_f: List[float] = cf
        self.min_size: int = spec.min_size          # defaults.py:59-60
        self.max_size: int = spec.max_size
        self.input_format: str = spec.input_format  # defaults.py:62
        # confidence heads of a WC* model (chart_with_confidence.py:50-89): extra output keys pred_densepose_<name>
        self.extra_names = [name for name, _ in spec.extra_heads]

    def forward(self, original_image: torch.Tensor, bgr: bool = True) -> Dict[str, torch.Tensor]:
        out = torch.ops.dpb200.forward(original_image, bgr, self.weights, self.table, self.names, self.cfg_i,
                                       self.cfg_f, self.dtype_probe)
        res = {
            "image_size": out[0],
            "pred_boxes": out[1],
            "scores": out[2],
            "pred_classes": out[3],
            "pred_densepose_coarse_segm": out[4],
            "pred_densepose_fine_segm": out[5],
            "pred_densepose_u": out[6],
            "pred_densepose_v": out[7],
        }
        for i in range(len(self.extra_names)):
            res["pred_densepose_" + self.extra_names[i]] = out[8 + i]
        return res
