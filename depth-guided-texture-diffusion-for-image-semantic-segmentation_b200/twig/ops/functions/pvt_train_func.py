"""Autograd Functions of the PVT-v2 blocks in training (cod.py:824-1002 under autograd).

Same convention as train_func.py (and the reference's ``MSDeformAttnFunction``, twig/ops/functions/
ms_deform_attn_func.py:19-46): one Function per module, ``forward`` saves what ``backward`` needs, ``backward`` is
``@once_differentiable`` and chains the gradient kernels of ``libdgtd_ops.so`` by hand.  The residual stream and every
gradient that crosses a Function boundary is fp32; inside, ``mode`` selects the operand type of the GEMMs
(fp32: exact CUDA-core kernels; bf16: tcgen05 forward / dgrad / wgrad with fp32 accumulation).

  PvtBlockFn    Block (cod.py:957-961): x + prompt -> norm1 -> Attention (q / kv / sr conv + norm / softmax / proj)
                -> DropPath residual -> norm2 -> Mlp (fc1 -> depthwise 3x3 -> GELU -> fc2) -> DropPath residual
  PatchEmbedFn  OverlapPatchEmbed (cod.py:995-1001): conv k/stride (pad k//2) -> nn.LayerNorm
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from .. import capi
from ..capi import BF16, F32, call, ptr, stream
from . import decoder_bank as DB
from . import pvt_func as PF
from . import texture_diffusion_func as OP
from . import train_func as TF


def _f32(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    return None if t is None else t.detach().contiguous().float()


def _as(t: torch.Tensor, mode: int) -> torch.Tensor:
    return t.detach().to(torch.bfloat16).contiguous() if mode == BF16 else t.detach().float().contiguous()


# ---- primitives ------------------------------------------------------------------------------------------------
def attention_bwd(q, kv, o, do, B, N, Nk, heads):
    """(dq, dkv) fp32 of softmax(q k^T / 8) v; q / kv / o in the forward's dtype, do fp32."""
    dq = torch.empty(q.shape, device=q.device, dtype=torch.float32)
    dkv = torch.empty(kv.shape, device=q.device, dtype=torch.float32)
    ws = torch.empty(capi.load().dgtd_attention_bwd_ws_floats(B, N, Nk, heads), device=q.device, dtype=torch.float32)
    call("dgtd_attention_bwd", ptr(q), ptr(kv), ptr(o), ptr(do), ptr(dq), ptr(dkv), ptr(ws), capi.dtype_code(q.dtype),
         B, N, Nk, heads, 64 ** -0.5, stream())
    return dq, dkv


def dwconv3_gelu_bwd(x, wT, bias, g):
    """x (B,h,w,C) fp32 | bf16, g fp32 -> (dx fp32, dwT (9,C), dbias (C))."""
    B, h, w, C = x.shape
    du = torch.empty(x.shape, device=x.device, dtype=torch.float32)
    dwT = torch.empty(9, C, device=x.device, dtype=torch.float32)
    db = torch.empty(C, device=x.device, dtype=torch.float32)
    ws = None
    if x.dtype == torch.bfloat16:
        ws = torch.empty(capi.load().dgtd_dwconv3_gelu_bwd_ws_floats(), device=x.device, dtype=torch.float32)
    call("dgtd_dwconv3_gelu_bwd", ptr(x), ptr(wT), ptr(bias), ptr(g), ptr(du), ptr(dwT), ptr(db), ptr(ws),
         capi.dtype_code(x.dtype), B, h, w, C, stream())
    dx = PF.dwconv3(du, wT.flip(0).contiguous(), torch.zeros_like(bias))      # rotated taps = input gradient
    return dx, dwT, db


def linear_bwd(g, a, w, mode, keep=None, rows_per_sample=1, need_da=True, da_out=None):
    """Gradients of out = keep * (a @ w^T + b) given g = dL/dout (fp32 (M, N)).

    a: (M, K) saved operand (bf16 in bf16 mode), w: (N, K) in the operand dtype.  Returns (da fp32 | None, dW fp32,
    db fp32).  bf16 mode runs dgrad / wgrad on tcgen05 from one fused cast + column-sum pass over g; small or ragged
    row counts (and fp32 mode) use the exact CUDA-core kernels.  `da_out`: accumulate da into this fp32 tensor."""
    N, K = w.shape
    M = g.numel() // N
    g = g.view(M, N)
    if mode == BF16 and TF.tc_rows_ok(M) and N % 8 == 0 and K % 8 == 0:
        gb, db = TF.eltwise_colsum(g, 0, keep=keep, rows_per_sample=rows_per_sample)
        dW = TF.wgrad_tc_mn(gb, a.view(M, K))
        da = None
        if need_da:
            da = OP.linear(gb, w.t().contiguous(), None, out_dtype=F32)
            if da_out is not None:
                da = da_out.add_(da.view(da_out.shape))
        return da, dW, db
    a32, w32 = a.view(M, K).float(), w.float()
    db = TF.colsum(g, N, keep, rows_per_sample)
    dW = TF.linear_wgrad(g, a32, N, K, keep, rows_per_sample)
    da = None
    if need_da:
        da = TF.linear_dgrad(g, w32, keep=keep, rows_per_sample=rows_per_sample,
                             out=None if da_out is None else da_out.view(M, K), accumulate=da_out is not None)
    return da, dW, db


def unpatchify_tokens(dp: torch.Tensor, B: int, h: int, w: int, C: int, sr: int) -> torch.Tensor:
    """Adjoint of `patchify_tokens` (non-overlapping patches: a pure permutation; data movement only)."""
    return dp.view(B, h // sr, w // sr, sr, sr, C).permute(0, 1, 3, 2, 4, 5).reshape(B, h, w, C)


# ---- Functions -------------------------------------------------------------------------------------------------
class PvtBlockFn(Function):
    """One PVT-v2 Block on the token stream.  cfg = (H, W, heads, sr, mode, eps_block, eps_sr)."""

    @staticmethod
    def forward(ctx, x, prompt, keep1, keep2, cfg, n1w, n1b, wq, bq, wkv, bkv, wsr, bsr, nsw, nsb, wp, bp, n2w, n2b,
                w1, b1, dww, dwb, w2, b2):
        H, W, heads, sr, mode, eps, eps_sr = cfg
        x = _f32(x)
        B, N, C = x.shape
        M = B * N
        n1w, n1b, n2w, n2b, bq, bkv, bp, b1, b2, dwb = map(_f32, (n1w, n1b, n2w, n2b, bq, bkv, bp, b1, b2, dwb))
        keep1, keep2 = _f32(keep1), _f32(keep2)
        wq_, wkv_, wp_, w1_, w2_ = (_as(t, mode) for t in (wq, wkv, wp, w1, w2))
        dwT = _f32(dww).reshape(-1, 9).t().contiguous()
        pr = None if prompt is None else prompt.detach().contiguous()
        a1, s = PF.ln_tokens(x, n1w, n1b, eps, mode, add=pr, want_sum=True)
        q = OP.linear(a1.view(M, C), wq_, bq)
        patches = red = wsr_ = None
        if sr > 1:
            wsr_ = _as(wsr.detach().permute(0, 2, 3, 1).reshape(C, sr * sr * C), mode)
            nsw, nsb, bsr = _f32(nsw), _f32(nsb), _f32(bsr)
            patches = PF.patchify_tokens(a1.view(B, H, W, C), sr)
            red = OP.linear(patches, wsr_, bsr, out_dtype=F32)
            xr, _ = PF.ln_tokens(red, nsw, nsb, eps_sr, mode)
            Nk = (H // sr) * (W // sr)
        else:
            xr, Nk = a1.view(M, C), N
        kv = OP.linear(xr, wkv_, bkv)
        o = PF.attention(q, kv, B, N, Nk, heads)
        x1 = torch.empty_like(s)
        OP.linear_residual_(o, wp_, bp, None, keep1, N, s, out=x1)
        a2, _ = PF.ln_tokens(x1, n2w, n2b, eps, mode)
        hpre = OP.linear(a2.view(M, C), w1_, b1)
        hid = hpre.shape[-1]
        gl = PF.dwconv3_gelu(hpre.view(B, H, W, hid), dwT, dwb)
        out = torch.empty_like(x1)
        OP.linear_residual_(gl.view(M, hid), w2_, b2, None, keep2, N, x1, out=out)
        ctx.cfg = cfg
        ctx.has_prompt = prompt is not None
        ctx.sr_shape = None if wsr is None else tuple(wsr.shape)
        ctx.dw_shape = tuple(dww.shape)
        ctx.save_for_backward(s, a1, q, patches, red, xr if sr > 1 else None, kv, o, x1, a2, hpre, gl, keep1, keep2,
                              n1w, wq_, wkv_, wsr_, nsw if sr > 1 else None, wp_, n2w, w1_, dwT, dwb, w2_)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        (s, a1, q, patches, red, xr, kv, o, x1, a2, hpre, gl, keep1, keep2, n1w, wq_, wkv_, wsr_, nsw, wp_, n2w, w1_,
         dwT, dwb, w2_) = ctx.saved_tensors
        H, W, heads, sr, mode, eps, eps_sr = ctx.cfg
        g = _f32(g)
        B, N, C = s.shape
        M = B * N
        hid = hpre.shape[-1]
        # Mlp branch: out = x1 + keep2 * (gl @ w2^T + b2)
        dgl, dW2, db2 = linear_bwd(g, gl, w2_, mode, keep2, N)
        dh, ddwT, ddwb = dwconv3_gelu_bwd(hpre.view(B, H, W, hid), dwT, dwb, dgl.view(B, H, W, hid))
        del dgl
        da2, dW1, db1 = linear_bwd(dh, a2, w1_, mode)
        del dh
        dx1, dn2w, dn2b = TF.ln_rows_bwd(da2.view(B, N, C), x1, n2w, eps)
        dx1.add_(g)
        # attention branch: x1 = s + keep1 * (o @ wp^T + bp)
        do, dWp, dbp = linear_bwd(dx1, o, wp_, mode, keep1, N)
        Nk = kv.shape[0] // B
        dq, dkv = attention_bwd(q, kv, o, do.view(M, C), B, N, Nk, heads)
        dWsr = dbsr = dnsw = dnsb = None
        if sr > 1:
            dxr, dWkv, dbkv = linear_bwd(dkv, xr, wkv_, mode)
            dred, dnsw, dnsb = TF.ln_rows_bwd(dxr, red, nsw, eps_sr)
            dpatch, dWsr_p, dbsr = linear_bwd(dred, patches, wsr_, mode)
            dWsr = dWsr_p.view(C, sr, sr, C).permute(0, 3, 1, 2).contiguous()
            da1 = unpatchify_tokens(dpatch, B, H, W, C, sr).contiguous().view(M, C)
            _, dWq, dbq = linear_bwd(dq, a1, wq_, mode, da_out=da1)
        else:
            da1, dWkv, dbkv = linear_bwd(dkv, a1, wkv_, mode)
            _, dWq, dbq = linear_bwd(dq, a1, wq_, mode, da_out=da1)
        ds, dn1w, dn1b = TF.ln_rows_bwd(da1.view(B, N, C), s, n1w, eps)
        ds.add_(dx1)
        grads = (ds, ds, None, None, None, dn1w, dn1b, dWq, dbq, dWkv, dbkv, dWsr, dbsr, dnsw, dnsb, dWp, dbp, dn2w, dn2b,
                 dW1, db1, ddwT.t().reshape(ctx.dw_shape), ddwb, dW2, db2)
        return tuple(gr if need else None for gr, need in zip(grads, ctx.needs_input_grad))


class PatchEmbedFn(Function):
    """OverlapPatchEmbed on an NHWC fp32 map: conv k x k / stride (pad k // 2) -> LayerNorm -> tokens (B, oh*ow, Cout).
    cfg = (k, stride, mode, eps)."""

    @staticmethod
    def forward(ctx, x, w, b, nw, nb, cfg):
        k, s, mode, eps = cfg
        x = _f32(x)
        B, H, W, Cin = x.shape
        Cout = w.shape[0]
        b, nw, nb = _f32(b), _f32(nw), _f32(nb)
        oh, ow = (H + 2 * (k // 2) - k) // s + 1, (W + 2 * (k // 2) - k) // s + 1
        unit = 8 if mode == BF16 else 4
        cp = (Cin + unit - 1) // unit * unit
        wpk = torch.zeros(Cout, k, k, cp, device=x.device, dtype=torch.float32)
        wpk[..., :Cin] = w.detach().float().permute(0, 2, 3, 1)
        wpk = _as(wpk.reshape(Cout, k * k * cp), mode)
        if cp != Cin:
            xp = torch.zeros(B, H, W, cp, device=x.device, dtype=torch.float32)
            xp[..., :Cin].copy_(x)
            x = xp
        if mode == BF16:
            col = DB.im2col(OP.cast(x, torch.bfloat16), k, s, -(k // 2), (oh, ow))
            pre = OP.linear(col, wpk, b, out_dtype=F32)
            saved_in = col
        else:
            pre = OP.conv_nhwc(x, wpk, b, cp, (oh, ow), k, s, -(k // 2))
            saved_in = x
        t, _ = PF.ln_tokens(pre.view(B, oh * ow, Cout), nw, nb, eps, F32)
        ctx.cfg = cfg
        ctx.geo = (B, H, W, Cin, cp, oh, ow, Cout)
        ctx.wshape = tuple(w.shape)
        ctx.save_for_backward(saved_in, pre, wpk, nw)
        return t

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        saved_in, pre, wpk, nw = ctx.saved_tensors
        k, s, mode, eps = ctx.cfg
        B, H, W, Cin, cp, oh, ow, Cout = ctx.geo
        M = B * oh * ow
        dpre, dnw, dnb = TF.ln_rows_bwd(_f32(g).view(M, Cout), pre.view(M, Cout), nw, eps)
        need_dx = ctx.needs_input_grad[0]
        if mode == BF16:
            dcol, dWp, db = linear_bwd(dpre, saved_in, wpk, mode, need_da=need_dx)
        else:
            db = TF.colsum(dpre, Cout)
            dWp = TF.linear_wgrad(dpre, saved_in, Cout, k * k * cp, conv=(k, H, W, cp, cp, oh, ow, s, -(k // 2)))
            dcol = TF.linear_dgrad(dpre, wpk) if need_dx else None
        dW = dWp.view(Cout, k, k, cp)[..., :Cin].permute(0, 3, 1, 2).contiguous()
        dx = None
        if need_dx:
            dxp = torch.empty(B, H, W, cp, device=g.device, dtype=torch.float32)
            DB.col2im(dcol, cp, None, dxp, cp, k, s, -(k // 2), (oh, ow))
            dx = dxp if cp == Cin else dxp[..., :Cin].contiguous()
        return dx, dW, db, dnw, dnb, None
