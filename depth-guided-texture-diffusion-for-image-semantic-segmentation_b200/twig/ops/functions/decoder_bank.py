"""Training path of the 16 ShapePropDecoders as ONE autograd Function (cod.py:1210-1226, 1308-1323,
and the prompt injection cod.py:1470-1505).

The forward is the inference design -- first convs of all decoders batched (they share the input),
last conv folded with the bilinear down-sample into a 4x4 stride-r conv so the
(B, E, 96, 96) decoder outputs of the reference are never materialised.  The backward keeps the same
geometry:

  conv3 (folded)  dW4 = im2col(h2)^T g  -> un-folded to the 3x3 weight;  dh2 = col2im(g W4) * relu'
  conv2           dW2 = im2col(h1)^T dh2;                                dh1 = conv(dh2, rot180(W2)^T) * relu'
  conv1 (batched) dW1 = im2col(emb)^T dh1;                               demb = sum_d conv(dh1_d, rot180(W1_d)^T)

precision "fp32": exact CUDA-core GEMMs (simt_gemm.cuh);  "bf16": tcgen05 grouped implicit-GEMM convs for
the forward and the input gradients, split-K tcgen05 GEMMs reading im2col / gradient matrices as MN-major
operands (no transposed copies) for the weight gradients.  All arithmetic is in libdgtd_ops.so; torch does allocation, views, packing copies.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from .. import capi
from ..capi import ACT_NONE, ACT_RELU, BF16, F32, call, ptr, stream
from . import texture_diffusion_func as OP
from . import train_func as TF

PAD = 32   # latent channels padded to one 64-byte swizzled row for the tcgen05 conv


# ---- weight packing (shared with the inference path) ------------------------------------------------------
def pack_conv3(wt: torch.Tensor) -> torch.Tensor:
    """(Cout,Cin,3,3) OIHW -> (Cout, 9*Cin) tap-major (dy,dx,c)."""
    return wt.detach().permute(0, 2, 3, 1).reshape(wt.shape[0], -1).float().contiguous()


def fold_conv3_bilinear(wt: torch.Tensor) -> torch.Tensor:
    """3x3 conv followed by the 2-tap (0.5/0.5 per axis) bilinear down-sample == 4x4 conv:
    W4[dy,dx] = 1/4 sum_{a,b in {0,1}} W3[dy-a, dx-b]    -> (Cout, 16*Cin)."""
    w3 = wt.detach().float()
    co, ci = w3.shape[0], w3.shape[1]
    w4 = torch.zeros(co, ci, 4, 4, device=w3.device, dtype=torch.float32)
    for a in (0, 1):
        for b in (0, 1):
            w4[:, :, a:a + 3, b:b + 3] += w3
    w4 *= 0.25
    return w4.permute(0, 2, 3, 1).reshape(co, -1).contiguous()


def unfold_grad4(dw4: torch.Tensor) -> torch.Tensor:
    """Adjoint of `fold_conv3_bilinear` on (Cout, 4, 4, Cin) -> (Cout, 3, 3, Cin)."""
    out = torch.zeros(dw4.shape[0], 3, 3, dw4.shape[3], device=dw4.device, dtype=dw4.dtype)
    for a in (0, 1):
        for b in (0, 1):
            out += dw4[:, a:a + 3, b:b + 3, :]
    return out * 0.25


def pad_taps_bf16(packed: torch.Tensor, cin: int, rows_to: Optional[int] = None) -> torch.Tensor:
    """(Cout, taps*cin) fp32 tap-major -> (rows_to or Cout, taps*32) bf16 with zero channel / row padding
    (operand layout of the tcgen05 implicit-GEMM conv: 32 channels = one 64-byte swizzled row)."""
    co = packed.shape[0]
    taps = packed.shape[1] // cin
    out = torch.zeros(rows_to or co, taps, PAD, device=packed.device, dtype=torch.float32)
    out[:co, :, :cin] = packed.reshape(co, taps, cin)
    return out.reshape(out.shape[0], taps * PAD).to(torch.bfloat16).contiguous()


def pad_rows(v: torch.Tensor, rows_to: int) -> torch.Tensor:
    out = torch.zeros(rows_to, device=v.device, dtype=torch.float32)
    out[:v.shape[0]] = v.detach().float()
    return out


def fold_params(src_hw: Tuple[int, int], dst_hw: Tuple[int, int]) -> Optional[Tuple[int, int]]:
    """(stride, offset) of the folded 4x4 conv when the bilinear resize src->dst is the exact
    2-tap average (integer ratio r in {2,4,8,...} on both axes); None otherwise."""
    (h, w), (oh, ow) = src_hw, dst_hw
    if oh <= 0 or ow <= 0 or h % oh or w % ow or h // oh != w // ow:
        return None
    r = h // oh
    if r < 2 or r % 2:
        return None
    return r, r // 2 - 2  # first averaged row is r*Y + r/2 - 1; the 3x3 conv reaches one row above


def conv_geometry(src_hw: Tuple[int, int], grid: Tuple[int, int]) -> Optional[Tuple[int, int, int]]:
    """(ks, stride, off) of the last decoder conv writing straight into the PVT token grid."""
    if tuple(src_hw) == tuple(grid):
        return 3, 1, -1
    f = fold_params(tuple(src_hw), tuple(grid))
    return None if f is None else (4, f[0], f[1])


def rot_t(wt: torch.Tensor) -> torch.Tensor:
    """Weights of the input-gradient conv: (Cout,Cin,3,3) -> (Cin, 9*Cout), channels transposed, taps
    rotated by 180 degrees."""
    return wt.flip(2, 3).permute(1, 2, 3, 0).reshape(wt.shape[1], -1).contiguous()


# ---- primitives -------------------------------------------------------------------------------------------
def col2im(dcol: torch.Tensor, Ct: int, mask: Optional[torch.Tensor], out: torch.Tensor, C: int,
           ks: int, stride: int, off: int, out_hw: Tuple[int, int]) -> torch.Tensor:
    """out (B,h,w,C slice, pixel pitch out.stride(2)) = relu-masked adjoint of im2col applied to dcol."""
    B, h, w = out.shape[0], out.shape[1], out.shape[2]
    ldc = dcol.shape[-1]
    call("dgtd_col2im_nhwc", dcol.data_ptr(), capi.dtype_code(dcol.dtype), ldc, Ct,
         mask.data_ptr() if mask is not None else None, mask.stride(2) if mask is not None else 0,
         out.data_ptr(), capi.dtype_code(out.dtype), out.stride(2), B, h, w, C, ks, stride, off,
         out_hw[0], out_hw[1], stream())
    return out


def im2col(x: torch.Tensor, ks: int, stride: int, off: int, out_hw: Tuple[int, int]) -> torch.Tensor:
    """x: (B,h,w,C) bf16 (C % 8 == 0; may be a channel slice, pixel pitch x.stride(2)) ->
    (B*oh*ow, ks*ks*C) bf16 tap-major, zero outside the map."""
    B, h, w, C = x.shape
    out = torch.empty(B * out_hw[0] * out_hw[1], ks * ks * C, device=x.device, dtype=torch.bfloat16)
    call("dgtd_im2col_nhwc", x.data_ptr(), x.stride(2), ptr(out), B, h, w, C, ks, stride, off, out_hw[0], out_hw[1],
         stream())
    return out


def group_sum(x: torch.Tensor, groups: int, group_stride: int, C: int) -> torch.Tensor:
    M = x.numel() // (groups * group_stride)
    out = torch.empty(M, C, device=x.device, dtype=torch.float32)
    call("dgtd_group_sum", ptr(x), capi.dtype_code(x.dtype), ptr(out), M, groups, group_stride, C, stream())
    return out


def bank_supported(src_hw: Tuple[int, int], grids: Sequence[Tuple[int, int]]) -> bool:
    return all(conv_geometry(src_hw, g) is not None for g in grids)


# ---- the Function ---------------------------------------------------------------------------------------------
class DecoderBankFn(Function):
    """apply(emb NHWC (B,h,w,L) fp32, cfg, *params) -> one (B, H_s*W_s, E_s) token tensor per decoder.

    cfg = {"stages": [(n_decoders, (H_s, W_s)), ...], "mode": F32 | BF16};  params = for every decoder
    (conv1.weight, conv1.bias, conv2.weight, conv2.bias, conv3.weight, conv3.bias) in stage order."""

    @staticmethod
    def forward(ctx, emb, cfg, *params):
        emb = TF._f32(emb)
        B, h, w, L = emb.shape
        D = len(params) // 6
        mode = cfg["mode"]
        P = [TF._f32(p) for p in params]
        geo = []                                  # per decoder: (ks, stride, off, grid)
        for n, grid in cfg["stages"]:
            g = conv_geometry((h, w), grid)
            assert g is not None, "DecoderBankFn needs an identity or power-of-two prompt injection"
            geo += [(g[0], g[1], g[2], tuple(grid))] * n
        assert len(geo) == D
        outs: List[torch.Tensor] = []
        if mode == BF16:
            emb_in = torch.zeros(B, h, w, PAD, device=emb.device, dtype=torch.bfloat16)
            emb_in[..., :L].copy_(emb)            # pad + cast (data movement)
            w1 = torch.cat([pad_taps_bf16(pack_conv3(P[6 * d]), L, PAD) for d in range(D)], 0)
            b1 = torch.cat([pad_rows(P[6 * d + 1], PAD) for d in range(D)])
            w2 = torch.cat([pad_taps_bf16(pack_conv3(P[6 * d + 2]), L, PAD) for d in range(D)], 0)
            b2 = torch.cat([pad_rows(P[6 * d + 3], PAD) for d in range(D)])
            h1 = torch.empty(B, h, w, PAD * D, device=emb.device, dtype=torch.bfloat16)
            OP.conv_nhwc_grouped(emb_in, w1, b1, PAD, (h, w), 3, 1, -1, ACT_RELU, h1, PAD * D, PAD * D, 1, 0, PAD * D, 0)
            h2 = torch.empty_like(h1)
            OP.conv_nhwc_grouped(h1, w2, b2, PAD, (h, w), 3, 1, -1, ACT_RELU, h2, PAD, PAD * D, D, PAD, PAD, PAD)
            d = 0
            for n, grid in cfg["stages"]:
                ks, st, off, _ = geo[d]
                E = P[6 * d + 4].shape[0]
                packer = pack_conv3 if ks == 3 else fold_conv3_bilinear
                w3 = torch.cat([pad_taps_bf16(packer(P[6 * (d + i) + 4]), L) for i in range(n)], 0)
                b3 = torch.cat([P[6 * (d + i) + 5] for i in range(n)]).contiguous()
                out = torch.empty(n, B, grid[0] * grid[1], E, device=emb.device, dtype=torch.bfloat16)
                OP.conv_nhwc_grouped(h2[..., d * PAD:], w3, b3, PAD, grid, ks, st, off, ACT_NONE, out, E, E, n, PAD, E,
                                     B * grid[0] * grid[1] * E)
                outs += [out[i] for i in range(n)]
                d += n
        else:
            emb_in = emb
            w1 = torch.cat([pack_conv3(P[6 * d]) for d in range(D)], 0).contiguous()
            b1 = torch.cat([P[6 * d + 1] for d in range(D)]).contiguous()
            h1 = OP.conv_nhwc(emb, w1, b1, L, (h, w), 3, 1, -1, act=ACT_RELU)
            h2 = torch.empty_like(h1)
            for d in range(D):
                OP.conv_nhwc(h1[..., d * L:(d + 1) * L], pack_conv3(P[6 * d + 2]), P[6 * d + 3], L, (h, w), 3, 1, -1,
                             act=ACT_RELU, out=h2[..., d * L:(d + 1) * L], Cout=L)
            for d in range(D):
                ks, st, off, grid = geo[d]
                w3 = pack_conv3(P[6 * d + 4]) if ks == 3 else fold_conv3_bilinear(P[6 * d + 4])
                y = OP.conv_nhwc(h2[..., d * L:(d + 1) * L], w3, P[6 * d + 5], L, grid, ks, st, off)
                outs.append(y.view(B, grid[0] * grid[1], -1))
        ctx.geo, ctx.mode, ctx.L = geo, mode, L
        ctx.save_for_backward(emb_in, h1, h2, *[P[6 * d + k] for d in range(D) for k in (0, 2, 4)])
        return tuple(outs)

    @staticmethod
    @once_differentiable
    def backward(ctx, *grads):
        emb_in, h1, h2, *W = ctx.saved_tensors
        D = len(W) // 3
        fn = _backward_bf16 if ctx.mode == BF16 else _backward_fp32
        demb, dparams = fn(emb_in, h1, h2, W, grads, ctx.geo, ctx.L, D)
        return (demb, None, *dparams)


def _unpack_w3_grad(dWp: torch.Tensor, E: int, ks: int, Ct: int, L: int) -> torch.Tensor:
    """(E, ks*ks*Ct) tap-major gradient of the (possibly folded) last conv -> (E, L, 3, 3)."""
    g4 = dWp.view(E, ks, ks, Ct)[..., :L]
    g3 = g4 if ks == 3 else unfold_grad4(g4)
    return g3.permute(0, 3, 1, 2).contiguous()


def _backward_fp32(emb, h1, h2, W, grads, geo, L, D):
    B, h, w, _ = emb.shape
    M = B * h * w
    dh1 = torch.empty(B, h, w, D * L, device=emb.device, dtype=torch.float32)
    dW2, db2, dW3, db3 = [], [], [], []
    for d in range(D):
        ks, st, off, grid = geo[d]
        w2_d, w3_d = W[3 * d + 1], W[3 * d + 2]
        E = w3_d.shape[0]
        g = TF._f32(grads[d]).reshape(-1, E)
        w3p = pack_conv3(w3_d) if ks == 3 else fold_conv3_bilinear(w3_d)
        h2_d, h1_d = h2[..., d * L:(d + 1) * L], h1[..., d * L:(d + 1) * L]
        dWp = TF.linear_wgrad(g, h2_d, E, ks * ks * L, conv=(ks, h, w, L, D * L, grid[0], grid[1], st, off))
        dW3.append(_unpack_w3_grad(dWp, E, ks, L, L))
        db3.append(TF.colsum(g, E))
        dcol = TF.linear_dgrad(g, w3p)                                        # (M_s, ks*ks*L)
        dh2_d = torch.empty(B, h, w, L, device=emb.device, dtype=torch.float32)
        col2im(dcol, L, h2_d, dh2_d, L, ks, st, off, grid)
        del dcol
        dWp2 = TF.linear_wgrad(dh2_d, h1_d, L, 9 * L, conv=(3, h, w, L, D * L, h, w, 1, -1))
        dW2.append(dWp2.view(L, 3, 3, L).permute(0, 3, 1, 2).contiguous())
        db2.append(TF.colsum(dh2_d, L))
        t = OP.conv_nhwc(dh2_d, rot_t(w2_d), None, L, (h, w), 3, 1, -1)
        col2im(t.view(M, L), L, h1_d, dh1[..., d * L:(d + 1) * L], L, 1, 1, 0, (h, w))
    dW1p = TF.linear_wgrad(dh1.view(M, D * L), emb, D * L, 9 * L, conv=(3, h, w, L, L, h, w, 1, -1))
    dW1 = dW1p.view(D, L, 3, 3, L).permute(0, 1, 4, 2, 3).contiguous()
    db1 = TF.colsum(dh1.view(M, D * L), D * L).view(D, L)
    w1_all = torch.cat([W[3 * d] for d in range(D)], 0)                      # (D*L, L, 3, 3)
    demb = OP.conv_nhwc(dh1, rot_t(w1_all), None, D * L, (h, w), 3, 1, -1)
    dparams = []
    for d in range(D):
        dparams += [dW1[d], db1[d].contiguous(), dW2[d], db2[d], dW3[d], db3[d]]
    return demb, dparams


def _backward_bf16(emb_pad, h1, h2, W, grads, geo, L, D):
    B, h, w, _ = emb_pad.shape
    M = B * h * w
    dev = emb_pad.device
    bf = torch.bfloat16
    dh2 = torch.empty(B, h, w, D * PAD, device=dev, dtype=bf)
    dW3, db3 = [], []
    for d in range(D):
        ks, st, off, grid = geo[d]
        w3_d = W[3 * d + 2]
        E = w3_d.shape[0]
        Ms = B * grid[0] * grid[1]
        g = grads[d].detach().reshape(Ms, E).to(bf).contiguous()
        h2_d = h2[..., d * PAD:(d + 1) * PAD]
        w3p = pad_taps_bf16(pack_conv3(w3_d) if ks == 3 else fold_conv3_bilinear(w3_d), L)   # (E, ks*ks*32)
        dWp = TF.wgrad_tc_mn(g, im2col(h2_d, ks, st, off, grid))                      # (E, ks*ks*32) = g^T col
        dW3.append(_unpack_w3_grad(dWp, E, ks, PAD, L))
        db3.append(TF.colsum_bf16(g))
        dcol = OP.linear(g, w3p.t().contiguous(), None)                                # (Ms, ks*ks*32) bf16
        col2im(dcol, PAD, h2_d, dh2[..., d * PAD:(d + 1) * PAD], PAD, ks, st, off, grid)
        del dcol
    # ---- conv2: per-decoder weight gradient, the decoder's 32 gradient channels read as a column slice
    dh2_2d = dh2.view(M, D * PAD)
    db2 = TF.colsum_bf16(dh2_2d).view(D, PAD)[:, :L]
    dW2 = []
    for d in range(D):
        o = TF.wgrad_tc_mn(dh2_2d[:, d * PAD:(d + 1) * PAD], im2col(h1[..., d * PAD:(d + 1) * PAD], 3, 1, -1, (h, w)))
        dW2.append(o.view(PAD, 3, 3, PAD)[:L, :, :, :L].permute(0, 3, 1, 2).contiguous())   # (co, ci, 3, 3)
    # ---- conv2 input gradient (grouped tcgen05 conv with transposed / rotated taps) + ReLU mask of h1
    w2t = torch.cat([pad_taps_bf16(rot_t(W[3 * d + 1]), L, PAD) for d in range(D)], 0)
    dh1 = torch.empty_like(dh2)
    OP.conv_nhwc_grouped(dh2, w2t, None, PAD, (h, w), 3, 1, -1, ACT_NONE, dh1, PAD, PAD * D, D, PAD, PAD, PAD)
    del dh2, dh2_2d
    col2im(dh1.view(M, D * PAD), D * PAD, h1, dh1, D * PAD, 1, 1, 0, (h, w))
    # ---- conv1: one weight-gradient GEMM for all decoders (shared input)
    dh1_2d = dh1.view(M, D * PAD)
    db1 = TF.colsum_bf16(dh1_2d).view(D, PAD)[:, :L]
    o = TF.wgrad_tc_mn(dh1_2d, im2col(emb_pad, 3, 1, -1, (h, w)))                         # (D*32, 288)
    dW1 = o.view(D, PAD, 3, 3, PAD)[:, :L, :, :, :L].permute(0, 1, 4, 2, 3).contiguous()
    # ---- input gradient of the shared embedding: per-decoder grouped conv, then the sum over decoders
    w1t = torch.cat([pad_taps_bf16(rot_t(W[3 * d]), L, PAD) for d in range(D)], 0)
    de = torch.empty(B, h, w, D * PAD, device=dev, dtype=bf)
    OP.conv_nhwc_grouped(dh1, w1t, None, PAD, (h, w), 3, 1, -1, ACT_NONE, de, PAD, PAD * D, D, PAD, PAD, PAD)
    demb = group_sum(de, D, PAD, L).view(B, h, w, L)
    dparams = []
    for d in range(D):
        dparams += [dW1[d], db1[d].contiguous(), dW2[d], db2[d].contiguous(), dW3[d], db3[d]]
    return demb, dparams
