"""Structure loss + deep supervision of the reference (cod.py:75-84, 135-141) on the CUDA path.

`cal_loss(preds, gts)` has the reference's signature and value; `boundary_weight(gts)` exposes the
label-only weight map so that a step computes it once for its five supervised maps
(`deep_supervision_loss`).  The SSIM term of cod.py:142-144 has no gradient path to any parameter
(embedding1 comes from the parameter-free FFT high-pass): `ssim_constant(embedding1, image)` returns its value
(forward only) and `total_loss` is the reference's `loss_p1 + loss_P2 + loss_3`.
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from ..ops import capi
from ..ops.capi import call, check_cuda, ptr, stream

__all__ = ["boundary_weight", "cal_loss", "deep_supervision_loss", "ssim_constant", "total_loss",
           "StructureLossFunction"]


def boundary_weight(gts: torch.Tensor) -> torch.Tensor:
    """1 + 5 |avgpool31x31(gts) - gts| (cod.py:76), gts (B,C,H,W) fp32."""
    gts = gts.detach().contiguous().float()
    check_cuda(gts)
    B, C, H, W = gts.shape
    out = torch.empty_like(gts)
    call("dgtd_boundary_weight_fwd", ptr(gts), ptr(out), B * C, H, W, stream())
    return out


class StructureLossFunction(Function):
    @staticmethod
    def forward(ctx, preds, gts, weit):
        preds = preds.contiguous().float()
        gts = gts.detach().contiguous().float()
        check_cuda(preds, gts, weit)
        B, C, H, W = preds.shape
        planes, HW = B * C, H * W
        ws = torch.empty(capi.load().dgtd_structure_loss_ws_floats(planes, HW), device=preds.device, dtype=torch.float32)
        sums = torch.empty(planes, 4, device=preds.device, dtype=torch.float32)
        loss = torch.empty(1, device=preds.device, dtype=torch.float32)
        call("dgtd_structure_loss_fwd", ptr(preds), ptr(gts), ptr(weit), ptr(ws), ptr(sums), ptr(loss), planes, HW, stream())
        ctx.save_for_backward(preds, gts, weit, sums)
        return loss.reshape(())

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_out):
        preds, gts, weit, sums = ctx.saved_tensors
        B, C, H, W = preds.shape
        g = grad_out.detach().reshape(1).contiguous().float()
        dp = torch.empty_like(preds)
        call("dgtd_structure_loss_bwd", ptr(preds), ptr(gts), ptr(weit), ptr(sums), ptr(g), ptr(dp), B * C, H * W, stream())
        return dp, None, None


def cal_loss(preds: torch.Tensor, gts: torch.Tensor, weit: Optional[torch.Tensor] = None) -> torch.Tensor:
    """`cod.cal_loss` (cod.py:75-84): weighted BCE + weighted IoU, mean over (image, channel) planes."""
    if weit is None:
        weit = boundary_weight(gts)
    return StructureLossFunction.apply(preds, gts, weit)


def deep_supervision_loss(P1: Sequence[torch.Tensor], P2: torch.Tensor, label: torch.Tensor, gamma: float = 0.2):
    """cod.py:135-141 without the constant SSIM term: sum_it (gamma * it) * cal_loss(P1[it]) + cal_loss(P2)."""
    weit = boundary_weight(label)
    loss = cal_loss(P2, label, weit)
    for it, out in enumerate(P1):
        if it > 0:                      # it = 0 carries weight 0 in the reference
            loss = loss + (gamma * it) * cal_loss(out, label, weit)
    return loss


@torch.no_grad()
def ssim_constant(embedding1: torch.Tensor, image: torch.Tensor) -> torch.Tensor:
    """cod.py:143-144: `ssim((e - e.min()) / (e.max() - e.min() + 1e-8), input)` with `SSIM` of cod.py:316-351.
    Value only: nothing trainable is upstream of embedding1."""
    e = embedding1.detach().contiguous().float()
    y = image.detach().contiguous().float()
    check_cuda(e, y)
    assert e.shape == y.shape and e.dim() == 4, (tuple(e.shape), tuple(y.shape))
    B, C, H, W = e.shape
    ws = torch.empty(capi.load().dgtd_ssim_loss_ws_floats(e.numel()), device=e.device, dtype=torch.float32)
    out = torch.empty(1, device=e.device, dtype=torch.float32)
    call("dgtd_ssim_loss_fwd", ptr(e), ptr(y), ptr(ws), ptr(out), B * C, H, W, stream())
    return out.reshape(())


def total_loss(embedding1: torch.Tensor, P1: Sequence[torch.Tensor], P2: torch.Tensor, image: torch.Tensor,
               label: torch.Tensor) -> torch.Tensor:
    """cod.py:135-146: loss_p1 + loss_P2 + loss_3 (the SSIM constant enters the value, not the gradient)."""
    return deep_supervision_loss(P1, P2, label) + ssim_constant(embedding1, image)
