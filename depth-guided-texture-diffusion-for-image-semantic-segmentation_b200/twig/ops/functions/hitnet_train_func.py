"""Autograd Functions of the Hitnet iterative decoder in training (SURVEY.md 8f-2 under autograd; cod.py:355-506,
685-807 with `cod.forward(mode='loss')`, cod.py:118-146).

Same convention as train_func.py / pvt_train_func.py: one Function per reference module, ``forward`` saves what
``backward`` needs, ``backward`` is ``@once_differentiable`` and chains the kernels of ``libdgtd_ops.so`` by hand
(csrc/hitnet_train.cu for what is specific to the decoder, the conv / GEMM gradient kernels of the trunk for the rest).
Maps are NHWC fp32 between Functions; ``mode`` selects the operand type of the convs (fp32: exact CUDA-core implicit
GEMMs; bf16: tcgen05 forward / dgrad / wgrad with fp32 accumulation, falling back to the exact kernels on maps too small
for a split-K tensor-core launch).

  ConvBnFn   BasicConv2d (cod.py:355-368): conv (no bias) -> BatchNorm2d, batch statistics in train(), running ones in eval()
  CabFn      CAB (cod.py:434-451): conv3 -> PReLU (ONE shared slope) -> conv3 -> CALayer -> + x
  SamFn      SAM (cod.py:454-506)
  Head1Fn    out_CFM / out_SAM (cod.py:710-711): 1-channel 1x1 conv with bias -> (B,1,h,w)
  ResizeLdFn nn.Upsample(mode='bilinear', align_corners=True) (cod.py:709,733,737) on NHWC
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from .. import capi
from ..capi import BF16, call, ptr, stream
from . import decoder_bank as DB
from . import hitnet_func as HF
from . import pvt_train_func as PT
from . import texture_diffusion_func as OP
from . import train_func as TF


def _f32(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    return None if t is None else t.detach().contiguous().float()


def _ws(nbytes: int, like: torch.Tensor) -> torch.Tensor:
    return torch.empty((nbytes + 7) // 8, device=like.device, dtype=torch.float64)


def _tap_major(w: torch.Tensor) -> torch.Tensor:
    """(O, I, kh, kw) -> (O, kh*kw*I) fp32: the K order of the implicit-GEMM loaders."""
    return w.detach().float().permute(0, 2, 3, 1).reshape(w.shape[0], -1).contiguous()


def _tap_major_padded_bf16(w: torch.Tensor) -> torch.Tensor:
    """(O, I, 3, 3) -> (O, 9 * Cp) bf16, Cp = roundup(I, 64) (the operand of dgtd_conv3x3_tc_fwd)."""
    O, I = w.shape[0], w.shape[1]
    Cp = (I + 63) // 64 * 64
    out = torch.zeros(O, 3, 3, Cp, device=w.device, dtype=torch.float32)
    out[..., :I] = w.detach().float().permute(0, 2, 3, 1)
    return out.reshape(O, 9 * Cp).to(torch.bfloat16).contiguous()


def _rot(w: torch.Tensor) -> torch.Tensor:
    """Weights of the input-gradient conv of a 3x3 / stride 1 / pad 1 conv: (O, I, 3, 3) -> (I, O, 3, 3), channels
    transposed, taps rotated by 180 degrees."""
    return w.detach().flip(2, 3).transpose(0, 1).contiguous()


# ---- conv without bias: forward / backward -----------------------------------------------------------------------
def conv_fwd(x: torch.Tensor, w: torch.Tensor, k: int, s: int, p: int, mode: int,
             prelu_in: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x (B,h,w,Cin) fp32 NHWC, w (Cout,Cin,k,k) -> conv(prelu_in(x)) (B,oh,ow,Cout) fp32."""
    B, h, wd, Cin = x.shape
    oh, ow = (h + 2 * p - k) // s + 1, (wd + 2 * p - k) // s + 1
    if mode == BF16:
        if (k, s, p) == (3, 1, 1):
            return HF.conv3_tc(x, _tap_major_padded_bf16(w), prelu_in=prelu_in)
        return HF.conv_affine_tc(x, _tap_major(w).to(torch.bfloat16), (oh, ow), k, s, -p, prelu_in=prelu_in)
    if prelu_in is not None:
        x = prelu_fwd(x, prelu_in)
    return HF.conv_affine(x, _tap_major(w), (oh, ow), k, s, -p)


def conv_bwd(g: torch.Tensor, x: torch.Tensor, w: torch.Tensor, k: int, s: int, p: int, mode: int,
             need_dx: bool = True) -> Tuple[Optional[torch.Tensor], torch.Tensor]:
    """g (B,oh,ow,Cout) fp32 dense, x the conv's input (B,h,w,Cin) fp32 -> (dx (B,h,w,Cin) fp32 | None, dW like w)."""
    B, h, wd, Cin = x.shape
    _, oh, ow, Cout = g.shape
    M = B * oh * ow
    K = k * k * Cin
    g2 = g.view(M, Cout)
    if mode == BF16 and TF.tc_rows_ok(M) and Cout % 8 == 0 and Cin % 8 == 0:
        if (k, s, p) == (3, 1, 1):
            Cp = (Cin + 63) // 64 * 64
            xb = torch.empty(B, h, wd, Cp, device=x.device, dtype=torch.bfloat16)
            call("dgtd_cast_pad_act_fwd", x.data_ptr(), Cin, ptr(xb), None, B * h * wd, Cin, Cp, stream())
            col = DB.im2col(xb, 3, 1, -1, (h, wd))
            dWp = TF.wgrad_tc_mn(OP.cast(g2, torch.bfloat16), col)
            dW = dWp.view(Cout, 3, 3, Cp)[..., :Cin].permute(0, 3, 1, 2).contiguous()
            dx = HF.conv3_tc(g, _tap_major_padded_bf16(_rot(w))) if need_dx else None
            return dx, dW
        col = x.view(M, Cin) if k == 1 and s == 1 else None
        col = OP.cast(col, torch.bfloat16) if col is not None else DB.im2col(OP.cast(x, torch.bfloat16), k, s, -p, (oh, ow))
        dcol, dWp, _ = PT.linear_bwd(g2, col, _tap_major(w).to(torch.bfloat16), mode, need_da=need_dx)
    else:
        wp = _tap_major(w)
        dWp = TF.linear_wgrad(g2, x, Cout, K, conv=(k, h, wd, Cin, Cin, oh, ow, s, -p))
        if not need_dx:
            dcol = None
        elif (k, s, p) == (3, 1, 1):
            return HF.conv_affine(g, _tap_major(_rot(w)), (h, wd), 3, 1, -1), \
                dWp.view(Cout, 3, 3, Cin).permute(0, 3, 1, 2).contiguous()
        else:
            dcol = TF.linear_dgrad(g2, wp)
    dW = dWp.view(Cout, k, k, Cin).permute(0, 3, 1, 2).contiguous()
    if not need_dx:
        return None, dW
    if k == 1 and s == 1:
        return dcol.view(B, h, wd, Cin), dW
    dx = torch.empty(B, h, wd, Cin, device=g.device, dtype=torch.float32)
    DB.col2im(dcol, Cin, None, dx, Cin, k, s, -p, (oh, ow))
    return dx, dW


# ---- primitives of csrc/hitnet_train.cu --------------------------------------------------------------------------
def prelu_fwd(u: torch.Tensor, slope: torch.Tensor) -> torch.Tensor:
    v = torch.empty_like(u)
    call("dgtd_prelu_fwd", ptr(u), ptr(slope), ptr(v), u.numel(), stream())
    return v


def prelu_bwd(u: torch.Tensor, g: torch.Tensor, slope: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    du = torch.empty_like(u)
    ds = torch.empty(1, device=u.device, dtype=torch.float32)
    ws = _ws(capi.load().dgtd_prelu_bwd_ws_bytes(u.numel()), u)
    call("dgtd_prelu_bwd", ptr(u), ptr(g), ptr(slope), ptr(du), ptr(ds), ptr(ws), u.numel(), stream())
    return du, ds


def bn_fwd(y: torch.Tensor, gamma, beta, run_mean, run_var, momentum: float, eps: float, batch_stats: bool):
    """y (.., C) fp32 dense -> (out, mean, rstd).  batch_stats: statistics of this batch (+ running update when the
    buffers are given); else the running statistics (eval-mode BatchNorm inside an autograd graph)."""
    C = y.shape[-1]
    M = y.numel() // C
    out = torch.empty_like(y)
    if batch_stats:
        mean = torch.empty(C, device=y.device, dtype=torch.float32)
        rstd = torch.empty(C, device=y.device, dtype=torch.float32)
        ws = _ws(capi.load().dgtd_col_stats_ws_bytes(M, C), y)
        call("dgtd_bn_train_fwd", ptr(y), C, ptr(gamma), ptr(beta), ptr(run_mean), ptr(run_var), float(momentum),
             float(eps), ptr(out), C, ptr(mean), ptr(rstd), ptr(ws), M, C, stream())
    else:
        mean = run_mean.detach().float().contiguous()
        rstd = torch.rsqrt(run_var.detach().float() + eps).contiguous()
        call("dgtd_bn_apply_fwd", ptr(y), C, ptr(mean), ptr(rstd), ptr(gamma), ptr(beta), ptr(out), C, M, C, stream())
    return out, mean, rstd


def bn_bwd(g: torch.Tensor, y: torch.Tensor, mean, rstd, gamma, batch_stats: bool):
    C = y.shape[-1]
    M = y.numel() // C
    dx = torch.empty_like(y)
    dgamma = torch.empty(C, device=y.device, dtype=torch.float32)
    dbeta = torch.empty(C, device=y.device, dtype=torch.float32)
    ws = _ws(capi.load().dgtd_col_stats_ws_bytes(M, C), y)
    call("dgtd_bn_train_bwd", ptr(g), C, ptr(y), C, ptr(mean), ptr(rstd), ptr(gamma), ptr(dx), C, ptr(dgamma), ptr(dbeta),
         ptr(ws), int(batch_stats), M, C, stream())
    return dx, dgamma, dbeta


def channel_dot(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """(B, chunks, C) fixed-order partial sums over the pixels of a * b."""
    B, h, w, C = a.shape
    nch = capi.load().dgtd_channel_sums_chunks(h * w)
    part = torch.empty(B, nch, C, device=a.device, dtype=torch.float32)
    call("dgtd_channel_dot_fwd", ptr(a), C, ptr(b), C, ptr(part), B, h * w, C, stream())
    return part


def gate_bwd(part, hw, dpart, w1, w2, v1=None, v2=None, grads=None):
    """-> (dmean (B,C), dw1, dw2, dv1, dv2); `grads` = previous (dw1, dw2, dv1, dv2) to accumulate into."""
    B, nch, C = part.shape
    Cr = w1.shape[0]
    Cs = 0 if v1 is None else v1.shape[0]
    dmean = torch.empty(B, C, device=part.device, dtype=torch.float32)
    if grads is None:
        dw1, dw2 = torch.empty_like(w1), torch.empty_like(w2)
        dv1 = torch.empty_like(v1) if Cs else None
        dv2 = torch.empty_like(v2) if Cs else None
    else:
        dw1, dw2, dv1, dv2 = grads
    ws = torch.empty(capi.load().dgtd_gate_bwd_ws_floats(B, C, Cr, Cs), device=part.device, dtype=torch.float32)
    call("dgtd_gate_bwd", ptr(part), nch, hw, ptr(dpart), dpart.shape[1], ptr(w1), ptr(w2), ptr(v1), ptr(v2), ptr(dmean),
         ptr(dw1), ptr(dw2), ptr(dv1), ptr(dv2), ptr(ws), int(grads is not None), B, C, Cr, Cs, stream())
    return dmean, dw1, dw2, dv1, dv2


def gated_bwd(g: torch.Tensor, gate: torch.Tensor, scal: Optional[torch.Tensor], dmean: Optional[torch.Tensor]):
    B, h, w, C = g.shape
    out = torch.empty_like(g)
    call("dgtd_gated_bwd", ptr(g), C, ptr(gate), ptr(scal), ptr(dmean), ptr(out), C, B, h * w, C, stream())
    return out


# ---- Functions ---------------------------------------------------------------------------------------------------
class ConvBnFn(Function):
    """cfg = (k, stride, pad, mode, eps, momentum, batch_stats)."""

    @staticmethod
    def forward(ctx, x, w, gamma, beta, run_mean, run_var, cfg):
        k, s, p, mode, eps, mom, batch_stats = cfg
        x = _f32(x)
        gamma, beta = _f32(gamma), _f32(beta)
        y = conv_fwd(x, w, k, s, p, mode)
        out, mean, rstd = bn_fwd(y, gamma, beta, run_mean, run_var, mom, eps, batch_stats)
        ctx.cfg = cfg
        ctx.save_for_backward(x, w.detach(), y, mean, rstd, gamma)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        x, w, y, mean, rstd, gamma = ctx.saved_tensors
        k, s, p, mode, eps, mom, batch_stats = ctx.cfg
        dy, dgamma, dbeta = bn_bwd(_f32(g), y, mean, rstd, gamma, batch_stats)
        dx, dW = conv_bwd(dy, x, w, k, s, p, mode, ctx.needs_input_grad[0])
        return dx, dW, dgamma, dbeta, None, None, None


class CabFn(Function):
    """out = CA(conv3(prelu(conv3(x)))) + x;  ca1 (Cr,C,1,1), ca2 (C,Cr,1,1)."""

    @staticmethod
    def forward(ctx, x, w0, slope, w2, ca1, ca2, mode):
        x = _f32(x)
        slope_ = _f32(slope).reshape(1)
        c1, c2 = _f32(ca1).flatten(1).contiguous(), _f32(ca2).flatten(1).contiguous()
        u = conv_fwd(x, w0, 3, 1, 1, mode)
        r = conv_fwd(u, w2, 3, 1, 1, mode, prelu_in=slope_)
        part, hw = HF.channel_sums(r)
        gate = HF.channel_gate(part, hw, c1, c2)
        out = HF.gated_sum(r, ga=gate, b=x)
        ctx.mode = mode
        ctx.shapes = (tuple(slope.shape), tuple(ca1.shape), tuple(ca2.shape))
        ctx.save_for_backward(x, w0.detach(), slope_, w2.detach(), c1, c2, u, r, part, gate)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        x, w0, slope_, w2, c1, c2, u, r, part, gate = ctx.saved_tensors
        mode = ctx.mode
        g = _f32(g)
        hw = r.shape[1] * r.shape[2]
        dmean, dc1, dc2, _, _ = gate_bwd(part, hw, channel_dot(g, r), c1, c2)
        dr = gated_bwd(g, gate, None, dmean)
        v = prelu_fwd(u, slope_)
        dv, dW2 = conv_bwd(dr, v, w2, 3, 1, 1, mode)
        del v, dr
        du, dslope = prelu_bwd(u, dv, slope_)
        dx, dW0 = conv_bwd(du, x, w0, 3, 1, 1, mode, ctx.needs_input_grad[0])
        if dx is not None:
            dx.add_(g)
        s_shape, c1_shape, c2_shape = ctx.shapes
        return dx, dW0, dslope.reshape(s_shape), dW2, dc1.reshape(c1_shape), dc2.reshape(c2_shape), None


class SamFn(Function):
    """x_h fc(mean x_h) fc_wight(mean x_h) + x_l fc(mean x_l) fc_wight(mean x_l);  f0 (Cr,C), f2 (C,Cr), g0 (Cs,C), g2 (1,Cs)."""

    @staticmethod
    def forward(ctx, x_h, x_l, f0, f2, g0, g2):
        x_h, x_l = _f32(x_h), _f32(x_l)
        f0_, f2_, g0_, g2_ = _f32(f0), _f32(f2), _f32(g0), _f32(g2)
        ph, hw_h = HF.channel_sums(x_h)
        pl, hw_l = HF.channel_sums(x_l)
        gh, sh = HF.channel_gate(ph, hw_h, f0_, f2_), HF.channel_gate(ph, hw_h, g0_, g2_)
        gl, sl = HF.channel_gate(pl, hw_l, f0_, f2_), HF.channel_gate(pl, hw_l, g0_, g2_)
        out = HF.gated_sum(x_h, ga=gh, sa=sh, b=x_l, gb=gl, sb=sl)
        ctx.save_for_backward(x_h, x_l, f0_, f2_, g0_, g2_, ph, pl, gh, sh, gl, sl)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        x_h, x_l, f0, f2, g0, g2, ph, pl, gh, sh, gl, sl = ctx.saved_tensors
        g = _f32(g)
        hw_h, hw_l = x_h.shape[1] * x_h.shape[2], x_l.shape[1] * x_l.shape[2]
        dm_h, df0, df2, dg0, dg2 = gate_bwd(ph, hw_h, channel_dot(g, x_h), f0, f2, g0, g2.reshape(-1))
        dm_l, df0, df2, dg0, dg2 = gate_bwd(pl, hw_l, channel_dot(g, x_l), f0, f2, g0, g2.reshape(-1),
                                            grads=(df0, df2, dg0, dg2))
        dx_h = gated_bwd(g, gh, sh.reshape(-1), dm_h)
        dx_l = gated_bwd(g, gl, sl.reshape(-1), dm_l)
        return dx_h, dx_l, df0, df2, dg0, dg2.reshape(g2.shape)


class Head1Fn(Function):
    """x (B,h,w,C) NHWC, w (1,C,1,1), b (1) -> (B,1,h,w)."""

    @staticmethod
    def forward(ctx, x, w, b):
        x = _f32(x)
        w_ = _f32(w).reshape(-1)
        ctx.wshape = tuple(w.shape)
        ctx.save_for_backward(x, w_)
        return HF.head1(x, w_, _f32(b))

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        x, w_ = ctx.saved_tensors
        g = _f32(g)
        B, h, wd, C = x.shape
        rows = B * h * wd
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        dw = torch.empty(C, device=x.device, dtype=torch.float32)
        db = torch.empty(1, device=x.device, dtype=torch.float32)
        ws = _ws(capi.load().dgtd_col_stats_ws_bytes(rows, C), x)
        call("dgtd_head1_bwd", ptr(g), ptr(x), C, ptr(w_), ptr(dx), C, ptr(dw), ptr(db), ptr(ws), rows, C, stream())
        return dx, dw.reshape(ctx.wshape), db


class ResizeLdFn(Function):
    """Bilinear resize of an NHWC fp32 map, either convention, with its adjoint."""

    @staticmethod
    def forward(ctx, x, size, align_corners):
        x = _f32(x)
        ctx.in_shape = tuple(x.shape)
        ctx.size = (int(size[0]), int(size[1]))
        ctx.align = bool(align_corners)
        return HF.resize_ld(x, ctx.size, ctx.align)

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        B, h, w, C = ctx.in_shape
        g = _f32(g)
        dx = torch.empty(ctx.in_shape, device=g.device, dtype=torch.float32)
        call("dgtd_resize_nhwc_ld_bwd", ptr(g), C, ptr(dx), C, B, h, w, C, ctx.size[0], ctx.size[1], int(ctx.align),
             stream())
        return dx, None, None
