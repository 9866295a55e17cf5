"""CPU restatement of the reference's structure loss and deep-supervision sum.

TEST INFRASTRUCTURE ONLY.  Follows `cod.cal_loss` (cod.py:75-84) and the combination in `cod.forward`
(cod.py:135-141); pinned against the unmodified reference by tests/golden/make_golden_loss.py.

  weit = 1 + 5 |avgpool31x31(gt) - gt|      (stride 1, zero padding 15, divisor always 961)
  wbce = sum(weit * bce_with_logits(p, gt)) / sum(weit)                       per (image, channel) plane
  wiou = 1 - (I + 1) / (U - I + 1),  I = sum(sigmoid(p) gt weit),  U = sum((sigmoid(p) + gt) weit)
  loss = mean over planes of (wbce + wiou)
  deep supervision: sum_it (0.2 * it) * loss(P1[it]) + loss(P2)     (it = 0 carries weight 0, cod.py:138-139)
"""
from __future__ import annotations

from typing import Sequence

import torch


def box_sum(x: torch.Tensor, k: int = 31) -> torch.Tensor:
    """Zero-padded k x k window sums via 2-D prefix sums (not avg_pool2d)."""
    r = k // 2
    H, W = x.shape[-2:]
    p = torch.zeros(x.shape[:-2] + (H + 2 * r + 1, W + 2 * r + 1), dtype=x.dtype)
    p[..., r + 1:r + 1 + H, r + 1:r + 1 + W] = x
    c = p.cumsum(-1).cumsum(-2)
    return c[..., k:, k:] - c[..., :-k, k:] - c[..., k:, :-k] + c[..., :-k, :-k]


def boundary_weight(gts: torch.Tensor) -> torch.Tensor:
    return 1 + 5 * (box_sum(gts, 31) / 961.0 - gts).abs()


def structure_loss(preds: torch.Tensor, gts: torch.Tensor) -> torch.Tensor:
    weit = boundary_weight(gts)
    bce = torch.clamp(preds, min=0) - preds * gts + torch.log1p(torch.exp(-preds.abs()))
    wbce = (weit * bce).sum(dim=(2, 3)) / weit.sum(dim=(2, 3))
    s = torch.sigmoid(preds)
    inter = (s * gts * weit).sum(dim=(2, 3))
    union = ((s + gts) * weit).sum(dim=(2, 3))
    wiou = 1 - (inter + 1) / (union - inter + 1)
    return (wbce + wiou).mean()


def deep_supervision_loss(P1: Sequence[torch.Tensor], P2: torch.Tensor, label: torch.Tensor, gamma: float = 0.2):
    loss = structure_loss(P2, label)
    for it, out in enumerate(P1):
        loss = loss + (gamma * it) * structure_loss(out, label)
    return loss


def ssim_constant(embedding1: torch.Tensor, image: torch.Tensor) -> torch.Tensor:
    """cod.py:143-144 + `SSIM._ssim` (cod.py:333-348): batch-wide min-max normalisation of embedding1, 3x3 means over
    reflection-padded windows written as nine shifted slices, mean of clamp((1 - SSIM) / 2, 0, 1)."""
    x = (embedding1 - embedding1.min()) / (embedding1.max() - embedding1.min() + 1e-8)
    y = image

    def mean3(t):
        H, W = t.shape[-2:]
        ry = [1] + list(range(H)) + [H - 2]
        rx = [1] + list(range(W)) + [W - 2]
        p = t[..., ry, :][..., :, rx]
        acc = torch.zeros_like(t)
        for dy in range(3):
            for dx in range(3):
                acc = acc + p[..., dy:dy + H, dx:dx + W]
        return acc / 9.0
    mx, my = mean3(x), mean3(y)
    vx, vy, vxy = mean3(x * x) - mx * mx, mean3(y * y) - my * my, mean3(x * y) - mx * my
    num = (2 * mx * my + 0.01 ** 2) * (2 * vxy + 0.03 ** 2)
    den = (mx * mx + my * my + 0.01 ** 2) * (vx + vy + 0.03 ** 2)
    return torch.clamp((1 - num / den) / 2, 0, 1).mean()


def total_loss(embedding1, P1, P2, image, label, gamma: float = 0.2):
    """cod.py:135-146."""
    return deep_supervision_loss(P1, P2, label, gamma) + ssim_constant(embedding1, image)
