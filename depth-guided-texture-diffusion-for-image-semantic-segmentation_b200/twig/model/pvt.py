"""PVT-v2 backbone with the texture prompts: mirror of cod.py:824-1002, 1340-1531, 1782-1787.

Same class names, constructor signatures, attribute names (identical ``state_dict`` keys) and return
values as the reference; every forward runs on libdgtd_ops.so.  Tokens are (B, N, C) = NHWC, the residual
stream is fp32, GEMM operands bf16 (tcgen05) or fp32 (exact mode), selected like the hot path
(``set_precision`` / ``torch.autocast``).  With gradients enabled and parameters that require them the same
forwards build an autograd graph through ``ops/functions/pvt_train_func.py`` (one Function per Block / patch embed,
backward on the same library); otherwise the fused inference kernels run.

``PyramidVisionTransformerImpr.forward_features`` is the fused pipeline: the hot path produces the prompt
tokens directly in token layout, `x + prompt[i]` (cod.py:1472) is fused into the block's first LayerNorm.
"""
from __future__ import annotations

import math
from functools import partial
from typing import List, Optional, Tuple

import torch
import torch.nn as nn

from ..ops.capi import BF16, F32
from ..ops.functions import pvt_func as PF
from ..ops.functions import pvt_train_func as PT
from ..ops.functions import texture_diffusion_func as OP
from ..ops.functions import train_func as TF
from . import texture_diffuser as TD
from .texture_diffuser import DropPath, _as, _mode, _packed, _wants_grad, prompt_decoder, prompt_encoder

__all__ = ["DWConv", "Mlp", "Attention", "Block", "OverlapPatchEmbed", "PyramidVisionTransformerImpr", "pvt_v2_b2"]


def _init_like_reference(m: nn.Module) -> None:
    """cod.py:1399-1414 (the same rule is repeated in every sub-module of the reference)."""
    if isinstance(m, nn.Linear):
        nn.init.trunc_normal_(m.weight, std=.02)
        if m.bias is not None:
            nn.init.constant_(m.bias, 0)
    elif isinstance(m, nn.LayerNorm):
        nn.init.constant_(m.bias, 0)
        nn.init.constant_(m.weight, 1.0)
    elif isinstance(m, nn.Conv2d):
        fan_out = m.kernel_size[0] * m.kernel_size[1] * m.out_channels // m.groups
        m.weight.data.normal_(0, math.sqrt(2.0 / fan_out))
        if m.bias is not None:
            m.bias.data.zero_()


def _f(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    return None if t is None else t.detach().float().contiguous()


class DWConv(nn.Module):
    """cod.py:1520-1531."""

    def __init__(self, dim=768):
        super(DWConv, self).__init__()
        self.dwconv = nn.Conv2d(dim, dim, 3, 1, 1, bias=True, groups=dim)

    def _packed_taps(self) -> torch.Tensor:
        return _packed(self).get("wT", [self.dwconv.weight],
                                 lambda: self.dwconv.weight.detach().reshape(-1, 9).t().float().contiguous())

    def forward(self, x, H, W):
        """(B, N, C) tokens -> depthwise 3x3 on the (H, W) grid -> (B, N, C), as cod.py:1526-1531 (inside `Mlp.forward`
        the same kernel runs fused with the GELU that follows)."""
        B, N, C = x.shape
        y = PF.dwconv3(x.contiguous().float().view(B, H, W, C), self._packed_taps(), _f(self.dwconv.bias))
        return y.view(B, N, C)


class Mlp(nn.Module):
    """cod.py:824-859."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.dwconv = DWConv(hidden_features)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features, out_features)
        self.drop = nn.Dropout(drop)
        self.apply(_init_like_reference)

    def _branch(self, a: torch.Tensor, H: int, W: int, mode: int, residual: Optional[torch.Tensor]):
        """a: (B, N, C) LayerNorm output in `mode` dtype; returns fc2(...) (+ residual, in place on it)."""
        pk = _packed(self)
        B, N, C = a.shape
        w1 = pk.get(f"fc1.{mode}", [self.fc1.weight], lambda: _as(self.fc1.weight, mode))
        w2 = pk.get(f"fc2.{mode}", [self.fc2.weight], lambda: _as(self.fc2.weight, mode))
        h = OP.linear(a.view(-1, C), w1, _f(self.fc1.bias))
        hid = h.shape[-1]
        g = PF.dwconv3_gelu(h.view(B, H, W, hid), self.dwconv._packed_taps(), _f(self.dwconv.dwconv.bias))
        if residual is None:
            return OP.linear(g.view(-1, hid), w2, _f(self.fc2.bias), out_dtype=F32).view(B, N, -1)
        OP.linear_residual_(g.view(-1, hid), w2, _f(self.fc2.bias), None, None, N, residual)
        return residual

    def forward(self, x, H, W):
        mode = _mode(self)
        return self._branch(_as(x, mode), H, W, mode, None)


class Attention(nn.Module):
    """cod.py:862-921 (head_dim must be 64, as in every pvt_v2 variant)."""

    def __init__(self, dim, num_heads=8, qkv_bias=False, qk_scale=None, attn_drop=0., proj_drop=0., sr_ratio=1):
        super().__init__()
        assert dim % num_heads == 0, f"dim {dim} should be divided by num_heads {num_heads}."
        self.dim = dim
        self.num_heads = num_heads
        head_dim = dim // num_heads
        self.scale = qk_scale or head_dim ** -0.5
        self.q = nn.Linear(dim, dim, bias=qkv_bias)
        self.kv = nn.Linear(dim, dim * 2, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)
        self.sr_ratio = sr_ratio
        if sr_ratio > 1:
            self.sr = nn.Conv2d(dim, dim, kernel_size=sr_ratio, stride=sr_ratio)
            self.norm = nn.LayerNorm(dim)
        self.apply(_init_like_reference)

    def _branch(self, a: torch.Tensor, H: int, W: int, mode: int, residual: Optional[torch.Tensor]):
        pk = _packed(self)
        B, N, C = a.shape
        assert C // self.num_heads == 64 and qk_scale_ok(self), "attention kernel is built for head_dim 64"
        wq = pk.get(f"q.{mode}", [self.q.weight], lambda: _as(self.q.weight, mode))
        wkv = pk.get(f"kv.{mode}", [self.kv.weight], lambda: _as(self.kv.weight, mode))
        wp = pk.get(f"proj.{mode}", [self.proj.weight], lambda: _as(self.proj.weight, mode))
        a2 = a.view(-1, C)
        q = OP.linear(a2, wq, _f(self.q.bias))
        if self.sr_ratio > 1:
            sr = self.sr_ratio
            wsr = pk.get(f"sr.{mode}", [self.sr.weight],
                         lambda: _as(self.sr.weight.detach().permute(0, 2, 3, 1).reshape(C, sr * sr * C), mode))
            patches = PF.patchify_tokens(a.view(B, H, W, C), sr)
            red = OP.linear(patches, wsr, _f(self.sr.bias), out_dtype=F32)
            xr, _ = PF.ln_tokens(red, _f(self.norm.weight), _f(self.norm.bias), self.norm.eps, mode)
            Nk = (H // sr) * (W // sr)
        else:
            xr, Nk = a2, N
        kv = OP.linear(xr, wkv, _f(self.kv.bias))
        o = PF.attention(q, kv, B, N, Nk, self.num_heads)
        if residual is None:
            return OP.linear(o, wp, _f(self.proj.bias), out_dtype=F32).view(B, N, C)
        OP.linear_residual_(o, wp, _f(self.proj.bias), None, None, N, residual)
        return residual

    def forward(self, x, H, W):
        mode = _mode(self)
        return self._branch(_as(x, mode), H, W, mode, None)


def qk_scale_ok(att: Attention) -> bool:
    return abs(att.scale - 64 ** -0.5) < 1e-12


class Block(nn.Module):
    """cod.py:924-961."""

    def __init__(self, dim, num_heads, mlp_ratio=4., qkv_bias=False, qk_scale=None, drop=0., attn_drop=0.,
                 drop_path=0., act_layer=nn.GELU, norm_layer=nn.LayerNorm, sr_ratio=1):
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.attn = Attention(dim, num_heads=num_heads, qkv_bias=qkv_bias, qk_scale=qk_scale, attn_drop=attn_drop,
                              proj_drop=drop, sr_ratio=sr_ratio)
        self.drop_path = DropPath(drop_path) if drop_path > 0. else nn.Identity()
        self.norm2 = norm_layer(dim)
        mlp_hidden_dim = int(dim * mlp_ratio)
        self.mlp = Mlp(in_features=dim, hidden_features=mlp_hidden_dim, act_layer=act_layer, drop=drop)
        self.apply(_init_like_reference)

    def _forward_tokens(self, x: torch.Tensor, prompt: Optional[torch.Tensor], H: int, W: int, mode: int) -> torch.Tensor:
        """x (B,N,C) fp32 residual stream (updated in place when no prompt is added), prompt (B,N,C) | None."""
        a, x = PF.ln_tokens(x, _f(self.norm1.weight), _f(self.norm1.bias), self.norm1.eps, mode, add=prompt, want_sum=True)
        x = self.attn._branch(a, H, W, mode, x)
        a, _ = PF.ln_tokens(x, _f(self.norm2.weight), _f(self.norm2.bias), self.norm2.eps, mode)
        return self.mlp._branch(a, H, W, mode, x)

    def _forward_train(self, x: torch.Tensor, prompt: Optional[torch.Tensor], H: int, W: int, mode: int) -> torch.Tensor:
        """Same block with an autograd graph (DropPath masks drawn per branch like timm's, cod.py:959-960)."""
        at, mlp = self.attn, self.mlp
        assert x.shape[-1] // at.num_heads == 64 and qk_scale_ok(at), "attention kernel is built for head_dim 64"
        dp = isinstance(self.drop_path, DropPath)
        keep1 = self.drop_path.keep_scale(x.shape[0], x.device) if dp else None
        keep2 = self.drop_path.keep_scale(x.shape[0], x.device) if dp else None
        sr = at.sr_ratio
        cfg = (H, W, at.num_heads, sr, mode, self.norm1.eps, at.norm.eps if sr > 1 else 0.0)
        return PT.PvtBlockFn.apply(
            x, prompt, keep1, keep2, cfg, self.norm1.weight, self.norm1.bias, at.q.weight, at.q.bias, at.kv.weight,
            at.kv.bias, at.sr.weight if sr > 1 else None, at.sr.bias if sr > 1 else None,
            at.norm.weight if sr > 1 else None, at.norm.bias if sr > 1 else None, at.proj.weight, at.proj.bias,
            self.norm2.weight, self.norm2.bias, mlp.fc1.weight, mlp.fc1.bias, mlp.dwconv.dwconv.weight,
            mlp.dwconv.dwconv.bias, mlp.fc2.weight, mlp.fc2.bias)

    def forward(self, x, H, W):
        if _wants_grad(self, x):
            return self._forward_train(x, None, H, W, _mode(self))
        return self._forward_tokens(x.detach().float().contiguous().clone(), None, H, W, _mode(self))


class OverlapPatchEmbed(nn.Module):
    """cod.py:964-1002."""

    def __init__(self, img_size=224, patch_size=7, stride=4, in_chans=3, embed_dim=768):
        super().__init__()
        img_size = (img_size, img_size) if isinstance(img_size, int) else tuple(img_size)
        patch_size = (patch_size, patch_size) if isinstance(patch_size, int) else tuple(patch_size)
        self.img_size = img_size
        self.patch_size = patch_size
        self.stride = stride
        self.H, self.W = img_size[0] // patch_size[0], img_size[1] // patch_size[1]
        self.num_patches = self.H * self.W
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=stride,
                              padding=(patch_size[0] // 2, patch_size[1] // 2))
        self.norm = nn.LayerNorm(embed_dim)
        self.apply(_init_like_reference)

    def _forward_nhwc(self, x: torch.Tensor, mode: int = F32) -> Tuple[torch.Tensor, int, int]:
        """x (B,H,W,Cin) fp32 -> tokens (B, H'*W', C) fp32.  Exact mode: CUDA-core implicit GEMM; bf16 mode:
        bf16 im2col (tap-major) + tcgen05 GEMM.  LayerNorm rows (eps of nn.LayerNorm) in fp32 either way."""
        B, H, W, Cin = x.shape
        k, s = self.patch_size[0], self.stride
        Cout = self.proj.out_channels
        unit = 8 if mode == BF16 else 4
        cp = (Cin + unit - 1) // unit * unit
        if cp != Cin or mode == BF16:             # RGB input: zero-pad the channels; bf16 mode: down-cast
            xp = torch.zeros(B, H, W, cp, device=x.device, dtype=torch.bfloat16 if mode == BF16 else torch.float32) \
                if cp != Cin else None
            if xp is None:
                x = OP.cast(x, torch.bfloat16)
            else:
                xp[..., :Cin].copy_(x)
                x = xp

        def pack():
            w = torch.zeros(Cout, k, k, cp, device=self.proj.weight.device, dtype=torch.float32)
            w[..., :Cin] = self.proj.weight.detach().float().permute(0, 2, 3, 1)
            return _as(w.reshape(Cout, k * k * cp), mode)
        wp = _packed(self).get(f"proj.{mode}.{cp}", [self.proj.weight], pack)
        oh, ow = (H + 2 * (k // 2) - k) // s + 1, (W + 2 * (k // 2) - k) // s + 1
        if mode == BF16:
            from ..ops.functions.decoder_bank import im2col
            y = OP.linear(im2col(x, k, s, -(k // 2), (oh, ow)), wp, _f(self.proj.bias), out_dtype=F32)
        else:
            y = OP.conv_nhwc(x, wp, _f(self.proj.bias), cp, (oh, ow), k, s, -(k // 2))
        t, _ = PF.ln_tokens(y.view(B, oh * ow, Cout), _f(self.norm.weight), _f(self.norm.bias), self.norm.eps, F32)
        return t, oh, ow

    def _forward_train(self, x: torch.Tensor, mode: int = F32) -> Tuple[torch.Tensor, int, int]:
        """x (B,H,W,Cin) fp32 NHWC -> tokens with an autograd graph (conv + LayerNorm as one Function)."""
        k, s = self.patch_size[0], self.stride
        t = PT.PatchEmbedFn.apply(x, self.proj.weight, self.proj.bias, self.norm.weight, self.norm.bias,
                                  (k, s, mode, self.norm.eps))
        H, W = x.shape[1], x.shape[2]
        return t, (H + 2 * (k // 2) - k) // s + 1, (W + 2 * (k // 2) - k) // s + 1

    def forward(self, x):
        if _wants_grad(self, x):
            return self._forward_train(TF.LayoutFn.apply(x, True), _mode(self))
        return self._forward_nhwc(OP.nchw_to_nhwc(x.detach().float().contiguous()), _mode(self))


class PyramidVisionTransformerImpr(nn.Module):
    """cod.py:1340-1509 (classification head / pretrained loading omitted: not on the path)."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, num_classes=1000, embed_dims=[64, 128, 256, 512],
                 num_heads=[1, 2, 4, 8], mlp_ratios=[4, 4, 4, 4], qkv_bias=False, qk_scale=None, drop_rate=0.,
                 attn_drop_rate=0., drop_path_rate=0., norm_layer=nn.LayerNorm,
                 depths=[3, 4, 6, 3], sr_ratios=[8, 4, 2, 1]):
        super().__init__()
        self.num_classes = num_classes
        self.depths = depths
        self.patch_embed1 = OverlapPatchEmbed(img_size=img_size, patch_size=7, stride=4, in_chans=in_chans,
                                              embed_dim=embed_dims[0])
        self.patch_embed2 = OverlapPatchEmbed(img_size=img_size // 4, patch_size=3, stride=2, in_chans=embed_dims[0],
                                              embed_dim=embed_dims[1])
        self.patch_embed3 = OverlapPatchEmbed(img_size=img_size // 8, patch_size=3, stride=2, in_chans=embed_dims[1],
                                              embed_dim=embed_dims[2])
        self.patch_embed4 = OverlapPatchEmbed(img_size=img_size // 16, patch_size=3, stride=2, in_chans=embed_dims[2],
                                              embed_dim=embed_dims[3])
        dpr = [x.item() for x in torch.linspace(0, drop_path_rate, sum(depths))]
        cur = 0
        for s in range(4):
            blocks = nn.ModuleList([Block(dim=embed_dims[s], num_heads=num_heads[s], mlp_ratio=mlp_ratios[s],
                                          qkv_bias=qkv_bias, qk_scale=qk_scale, drop=drop_rate, attn_drop=attn_drop_rate,
                                          drop_path=dpr[cur + i], norm_layer=norm_layer, sr_ratio=sr_ratios[s])
                                    for i in range(depths[s])])
            setattr(self, f"block{s + 1}", blocks)
            setattr(self, f"norm{s + 1}", norm_layer(embed_dims[s]))
            cur += depths[s]
        self.latent_dim = 24
        self.prompt_encoder = prompt_encoder(self.latent_dim, embed_dims, depths, True)
        self.prompt_decoder = nn.Sequential(*[prompt_decoder(self.latent_dim, embed_dims[i], depths[i], True)
                                              for i in range(len(depths))])
        self.apply(_init_like_reference)
        self.batch = 0

    @torch.no_grad()
    def _forward_features_nhwc(self, x, depth):
        """cod.py:1455-1509 with the stage maps left in NHWC (B, H_s, W_s, C_s) fp32 for the decoder kernels."""
        self.batch += 1
        if isinstance(depth, (list, tuple)):
            depth = torch.stack([d.reshape(1, *d.shape[-2:]) for d in depth], 0)
        mode = _mode(self)
        image = x.detach().float().contiguous()
        emb1, _, tokens = TD.texture_prompts(self.prompt_encoder, self.prompt_decoder, image, depth,
                                             precision="bf16" if mode == BF16 else "fp32", want_embedding3=False)
        B = image.shape[0]
        outs: List[torch.Tensor] = []
        cur = OP.nchw_to_nhwc(image)
        for s in range(4):
            t, H, W = getattr(self, f"patch_embed{s + 1}")._forward_nhwc(cur, mode)
            for i, blk in enumerate(getattr(self, f"block{s + 1}")):
                t = blk._forward_tokens(t, tokens[s][i].reshape(t.shape), H, W, mode)
            norm = getattr(self, f"norm{s + 1}")
            o, _ = PF.ln_tokens(t, _f(norm.weight), _f(norm.bias), norm.eps, F32)
            cur = o.view(B, H, W, -1)
            outs.append(cur)
        return emb1, outs

    def _forward_features_train(self, x, depth):
        """cod.py:1455-1509 with an autograd graph: hot path through `texture_prompts_train`, blocks and patch embeds
        through PvtBlockFn / PatchEmbedFn; stage maps stay NHWC fp32."""
        self.batch += 1
        if isinstance(depth, (list, tuple)):
            depth = torch.stack([d.reshape(1, *d.shape[-2:]) for d in depth], 0)
        mode = _mode(self)
        image = x.float().contiguous()
        emb1, _, tokens = TD.texture_prompts_train(self.prompt_encoder, self.prompt_decoder, image, depth,
                                                   precision="bf16" if mode == BF16 else "fp32")
        B = image.shape[0]
        outs: List[torch.Tensor] = []
        cur = TF.LayoutFn.apply(image, True) if image.requires_grad else OP.nchw_to_nhwc(image.detach())
        for s in range(4):
            t, H, W = getattr(self, f"patch_embed{s + 1}")._forward_train(cur, mode)
            for i, blk in enumerate(getattr(self, f"block{s + 1}")):
                t = blk._forward_train(t, tokens[s][i].reshape(t.shape), H, W, mode)
            norm = getattr(self, f"norm{s + 1}")
            cur = TF.LayerNormRowsFn.apply(t, norm.weight, norm.bias, norm.eps).view(B, H, W, -1)
            outs.append(cur)
        return emb1, outs

    def forward_features(self, x, depth):
        """cod.py:1455-1509 -> (embedding1, [(B, C_s, H_s, W_s) fp32]).  `depth` may be the reference's list of
        (1,H,W) maps or a (B,1,H,W) tensor.  Builds an autograd graph when gradients are wanted."""
        if _wants_grad(self, x):
            emb1, outs = self._forward_features_train(x, depth)
            return emb1, [TF.LayoutFn.apply(o, False) for o in outs]
        with torch.no_grad():
            emb1, outs = self._forward_features_nhwc(x, depth)
            return emb1, [OP.nhwc_to_nchw(o) for o in outs]

    def forward(self, x, depth):
        return self.forward_features(x, depth)


class pvt_v2_b2(PyramidVisionTransformerImpr):
    """cod.py:1782-1787."""

    def __init__(self, **kwargs):
        super(pvt_v2_b2, self).__init__(
            patch_size=4, embed_dims=[64, 128, 320, 512], num_heads=[1, 2, 5, 8], mlp_ratios=[8, 8, 4, 4],
            qkv_bias=True, norm_layer=partial(nn.LayerNorm, eps=1e-6), depths=[3, 4, 6, 3], sr_ratios=[8, 4, 2, 1],
            drop_rate=0.0, drop_path_rate=0.1)
