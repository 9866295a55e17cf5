"""Operator functions of the PVT-v2 blocks that consume the texture prompts: Python side
of csrc/pvt_ops.cu, same conventions as texture_diffusion_func.py (torch allocates, the library computes)."""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from .. import capi
from ..capi import BF16, F32, call, check_cuda, ptr, stream

__all__ = ["ln_tokens", "patchify_tokens", "dwconv3_gelu", "attention"]


def _tdtype(code: int) -> torch.dtype:
    return torch.float32 if code == F32 else torch.bfloat16


def ln_tokens(x: torch.Tensor, w, b, eps: float, out_dtype: int, add: Optional[torch.Tensor] = None,
              want_sum: bool = False) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """(LayerNorm_C(x + add), x + add | None); x fp32 (..., C), add fp32 | bf16 of the same shape."""
    check_cuda(x, w, b, add)
    assert x.dtype == torch.float32
    C = x.shape[-1]
    rows = x.numel() // C
    out = torch.empty(x.shape, device=x.device, dtype=_tdtype(out_dtype))
    s = torch.empty_like(x) if (want_sum and add is not None) else None
    call("dgtd_ln_tokens_fwd", ptr(x), ptr(add), capi.dtype_code(add.dtype) if add is not None else F32, ptr(s),
         ptr(w), ptr(b), ptr(out), out_dtype, rows, C, float(eps), stream())
    return out, (s if s is not None else (x if want_sum else None))


def patchify_tokens(x: torch.Tensor, sr: int) -> torch.Tensor:
    """x (B,h,w,C) -> (B*(h/sr)*(w/sr), sr*sr*C), tap-major: the operand of the spatial-reduction conv GEMM."""
    check_cuda(x)
    B, h, w, C = x.shape
    out = torch.empty(B * (h // sr) * (w // sr), sr * sr * C, device=x.device, dtype=x.dtype)
    call("dgtd_patchify_tokens_fwd", ptr(x), ptr(out), capi.dtype_code(x.dtype), B, h, w, C, sr, stream())
    return out


def dwconv3_gelu(x: torch.Tensor, wT: torch.Tensor, bias: torch.Tensor) -> torch.Tensor:
    """x (B,h,w,C) tokens (fp32 | bf16), wT (9,C) fp32: GELU(depthwise 3x3 pad 1 + bias)."""
    check_cuda(x, wT, bias)
    B, h, w, C = x.shape
    out = torch.empty_like(x)
    call("dgtd_dwconv3_gelu_fwd", ptr(x), ptr(wT), ptr(bias), ptr(out), capi.dtype_code(x.dtype), B, h, w, C, stream())
    return out


def dwconv3(x: torch.Tensor, wT: torch.Tensor, bias: torch.Tensor) -> torch.Tensor:
    """x (B,h,w,C) tokens (fp32 | bf16), wT (9,C) fp32: depthwise 3x3 pad 1 + bias (no activation)."""
    check_cuda(x, wT, bias)
    B, h, w, C = x.shape
    out = torch.empty_like(x)
    call("dgtd_dwconv3_fwd", ptr(x), ptr(wT), ptr(bias), ptr(out), capi.dtype_code(x.dtype), B, h, w, C, stream())
    return out


def attention(q: torch.Tensor, kv: torch.Tensor, B: int, N: int, Nk: int, heads: int) -> torch.Tensor:
    """q (B*N, heads*64), kv (B*Nk, 2*heads*64) -> softmax(q k^T / 8) v as (B*N, heads*64)."""
    check_cuda(q, kv)
    assert q.dtype == kv.dtype and q.shape[-1] == heads * 64 and kv.shape[-1] == 2 * heads * 64
    out = torch.empty_like(q)
    call("dgtd_attention_fwd", ptr(q), ptr(kv), ptr(out), capi.dtype_code(q.dtype), B, N, Nk, heads, 64 ** -0.5, stream())
    return out
