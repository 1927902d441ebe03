"""Helpers shared by the golden-fixture tests (same definitions as tests/golden/make_golden.py)."""
import torch


def label_map(out, n):
    """Native-resolution (4S x 4S) part labels of the first n detections: argmax25(fine) * (argmaxK(coarse) > 0)."""
    fine = out["pred_densepose_fine_segm"][:n].argmax(1)
    fg = out["pred_densepose_coarse_segm"][:n].argmax(1) > 0
    return (fine * fg).to(torch.uint8).cpu()
