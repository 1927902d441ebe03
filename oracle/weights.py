"""TEST INFRASTRUCTURE: the data-dependent calibration pass behind the synthetic weights.

The seeded weight / image generators themselves live in densepose_torchscript_b200/synth.py (the bench
and smoke paths need them without importing the oracle); they are re-exported here for the tests.
`calibrate` runs the fp32 oracle once per config and records one scalar per conv/linear layer
(pre-norm output std -> target); `python -m oracle.weights --calibrate` rewrites
densepose_torchscript_b200/synth_weight_scales.json.
"""
import json
import math
import os
from typing import Dict, List, Optional, Tuple

import torch

from densepose_torchscript_b200 import synth as _synth
from densepose_torchscript_b200.synth import (SCALES_PATH, TARGETS, add_aliases, layer_table, load_scales,  # noqa: F401
                                              make_state_dict, synthetic_image)

from . import densepose_oracle as O


def calibrate(spec: O.ModelSpec, seed: int = 0, height: int = 800, width: int = 1333) -> Dict[str, float]:
    """One forward pass of the fp32 oracle; each conv/linear weight is rescaled (in execution order) so its
    pre-norm output std equals its target. Returns {prefix: scale} rounded to 6 significant digits."""
    sd = make_state_dict(spec, seed, scales={}, use_committed_scales=False)
    scales: Dict[str, float] = {}

    def hook(sd_, prefix, fn):
        if prefix in scales:            # shared RPN head: calibrate on the first (largest) level only
            return
        y = fn()
        if y.numel() < 2:
            return
        std = float(y.std())
        s = float(f"{TARGETS.get(prefix, 1.0) / max(std, 1e-12):.6g}")
        sd_[prefix + ".weight"] *= s
        scales[prefix] = s

    O.CALIBRATOR = hook
    try:
        out = O.forward(synthetic_image(height, width, seed=1), sd, spec, mode="ref")
    finally:
        O.CALIBRATOR = None
    print(spec.name, "calibrated; detections:", len(out["scores"]))
    return scales


if __name__ == "__main__":
    import argparse

    ap = argparse.ArgumentParser()
    ap.add_argument("--calibrate", action="store_true")
    ap.add_argument("--configs", nargs="*", default=list(O.SPECS))
    a = ap.parse_args()
    if a.calibrate:
        allsc = {}
        if os.path.exists(SCALES_PATH):
            allsc = json.load(open(SCALES_PATH))
        for name in a.configs:
            allsc[name] = calibrate(O.SPECS[name])
        with open(SCALES_PATH, "w") as f:
            json.dump(allsc, f, indent=0, sort_keys=True)
        print("wrote", SCALES_PATH)
