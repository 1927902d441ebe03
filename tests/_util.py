"""Shared helpers for the parity tests."""
import torch


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.float().cpu(), b.float().cpu()
    assert a.shape == b.shape, (tuple(a.shape), tuple(b.shape))
    return float((a - b).norm() / max(float(b.norm()), 1e-12))


def frac_equal(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.float().cpu(), b.float().cpu()
    return float((a == b).float().mean())


def bf16(x: torch.Tensor) -> torch.Tensor:
    return x.to(torch.bfloat16).to(torch.float32)


def nchw(t: torch.Tensor) -> torch.Tensor:
    """engine NHWC tensor -> NCHW float cpu"""
    return t.permute(0, 3, 1, 2).float().cpu()


def nhwc_bf16_cuda(t: torch.Tensor) -> torch.Tensor:
    """NCHW float cpu -> NHWC bf16 cuda contiguous"""
    return t.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).cuda()


def box_iou(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    a, b = a.float().cpu(), b.float().cpu()
    area_a = (a[:, 2] - a[:, 0]).clamp(min=0) * (a[:, 3] - a[:, 1]).clamp(min=0)
    area_b = (b[:, 2] - b[:, 0]).clamp(min=0) * (b[:, 3] - b[:, 1]).clamp(min=0)
    lt = torch.maximum(a[:, None, :2], b[None, :, :2])
    rb = torch.minimum(a[:, None, 2:], b[None, :, 2:])
    wh = (rb - lt).clamp(min=0)
    inter = wh[..., 0] * wh[..., 1]
    return inter / (area_a[:, None] + area_b[None, :] - inter).clamp(min=1e-9)


def coord_dist(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """[Na, Nb] max-abs corner distance in pixels. (IoU is useless here: quirks 1 and 3 legitimately produce
    zero-width boxes, whose IoU with an identical box is 0/0.)"""
    a, b = a.float().cpu(), b.float().cpu()
    return (a[:, None, :] - b[None, :, :]).abs().amax(dim=2)


def match_detections(boxes_a: torch.Tensor, boxes_b: torch.Tensor, tol_px: float = 2.0):
    """Greedy one-to-one matching by corner distance. Returns (idx_a, idx_b) of matched pairs."""
    if len(boxes_a) == 0 or len(boxes_b) == 0:
        return torch.empty(0, dtype=torch.long), torch.empty(0, dtype=torch.long)
    dist = coord_dist(boxes_a, boxes_b)
    ia, ib = [], []
    for i in range(dist.shape[0]):
        j = int(dist[i].argmin())
        if dist[i, j] <= tol_px:
            ia.append(i); ib.append(j)
            dist[:, j] = float("inf")
    return torch.tensor(ia, dtype=torch.long), torch.tensor(ib, dtype=torch.long)
