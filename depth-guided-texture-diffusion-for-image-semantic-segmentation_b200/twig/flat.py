"""Layout of the flat fp32 parameter / gradient / moment buffers shared by `graphs.GraphedTrainStep`, the bucketed
gradient all-reduce and `optim.FusedAdamW`: parameters in registration order, every parameter starting on a
128-byte boundary (32 floats).  The kernels read parameters with float4 / TMA accesses, so a tensor re-homed into a
flat buffer must keep 16-byte alignment whatever the sizes of the tensors before it (PReLU slopes, 1-channel biases)."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch

ALIGN = 32   # elements


def flat_offsets(params: Sequence[torch.Tensor], align: int = ALIGN) -> Tuple[List[int], int]:
    """(offset of every parameter, total elements incl. padding)."""
    offs, off = [], 0
    for p in params:
        offs.append(off)
        off += (p.numel() + align - 1) // align * align
    return offs, off


def bind_views(params: Sequence[torch.nn.Parameter], flat: torch.Tensor, attr: str) -> List[int]:
    """Make `p.grad` (attr == "grad") or `p.data` (attr == "data", values preserved) of every parameter a view
    of `flat` at the layout's offsets; returns the offsets."""
    offs, total = flat_offsets(params)
    assert flat.numel() == total and flat.dtype == torch.float32, (flat.numel(), total)
    with torch.no_grad():
        for p, off in zip(params, offs):
            view = flat[off:off + p.numel()].view_as(p)
            if attr == "grad":
                p.grad = view
            else:
                view.copy_(p.detach())
                p.data = view
    return offs


def views_match(params: Sequence[torch.nn.Parameter], flat: torch.Tensor) -> bool:
    offs, total = flat_offsets(params)
    if flat.numel() != total:
        return False
    return all(p.grad is not None and p.grad.data_ptr() == flat.data_ptr() + 4 * off for p, off in zip(params, offs))
