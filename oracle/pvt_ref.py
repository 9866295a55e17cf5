"""CPU restatement of the PVT-v2 backbone that consumes the texture prompts.

TEST INFRASTRUCTURE ONLY: imported by tests/, tests/golden/make_golden_pvt.py and bench.py's CPU legs;
never by the product path.  Plain functional torch on a state dict with the reference's keys
(`PyramidVisionTransformerImpr`, cod.py:1340-1509); pinned against the unmodified reference classes by
tests/golden/make_golden_pvt.py (fixture tests/golden/pvt_*.npz).

Reference lines followed:
  OverlapPatchEmbed   cod.py:964-1002   conv k/stride, pad k//2 -> flatten -> nn.LayerNorm (eps 1e-5, the default)
  Attention           cod.py:862-921    q / kv Linear, spatial-reduction conv (k = stride = sr) + nn.LayerNorm
                                        (eps 1e-5) when sr > 1, softmax(q k^T * head_dim^-0.5) v, proj
  DWConv / Mlp        cod.py:1520-1531, 824-859   fc1 -> depthwise 3x3 (pad 1) -> GELU(erf) -> fc2
  Block               cod.py:924-961    x + attn(norm1(x)); x + mlp(norm2(x))   (norm eps 1e-6, pvt_v2_b2 :1782-1787)
  forward_features    cod.py:1455-1509  per stage: patch embed; for block i: x = blk(x + prompt[i]); norm; NCHW
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple

import torch
import torch.nn.functional as F

from . import texture_diffuser_ref as TDR

Params = Dict[str, torch.Tensor]

EMBED_DIMS = (64, 128, 320, 512)
NUM_HEADS = (1, 2, 5, 8)
MLP_RATIOS = (8, 8, 4, 4)
DEPTHS = (3, 4, 6, 3)
SR_RATIOS = (8, 4, 2, 1)
BLOCK_EPS = 1e-6       # norm_layer=partial(nn.LayerNorm, eps=1e-6), cod.py:1786
DEFAULT_EPS = 1e-5     # nn.LayerNorm(dim) inside OverlapPatchEmbed / Attention


def sub(p: Params, prefix: str) -> Params:
    pre = prefix + "."
    return {k[len(pre):]: v for k, v in p.items() if k.startswith(pre)}


def overlap_patch_embed(x: torch.Tensor, p: Params, patch: int, stride: int) -> Tuple[torch.Tensor, int, int]:
    """cod.py:995-1001: (B,Cin,H,W) -> tokens (B, H'*W', C), H', W'."""
    y = F.conv2d(x, p["proj.weight"], p["proj.bias"], stride=stride, padding=patch // 2)
    B, C, H, W = y.shape
    y = y.flatten(2).transpose(1, 2)
    y = F.layer_norm(y, (C,), p["norm.weight"], p["norm.bias"], DEFAULT_EPS)
    return y, H, W


def attention(x: torch.Tensor, H: int, W: int, p: Params, heads: int, sr: int) -> torch.Tensor:
    """cod.py:898-921."""
    B, N, C = x.shape
    d = C // heads
    q = F.linear(x, p["q.weight"], p.get("q.bias")).reshape(B, N, heads, d).permute(0, 2, 1, 3)
    if sr > 1:
        x_ = x.permute(0, 2, 1).reshape(B, C, H, W)
        x_ = F.conv2d(x_, p["sr.weight"], p["sr.bias"], stride=sr).reshape(B, C, -1).permute(0, 2, 1)
        x_ = F.layer_norm(x_, (C,), p["norm.weight"], p["norm.bias"], DEFAULT_EPS)
    else:
        x_ = x
    kv = F.linear(x_, p["kv.weight"], p.get("kv.bias")).reshape(B, -1, 2, heads, d).permute(2, 0, 3, 1, 4)
    k, v = kv[0], kv[1]
    attn = (q @ k.transpose(-2, -1)) * (d ** -0.5)
    attn = attn.softmax(dim=-1)
    y = (attn @ v).transpose(1, 2).reshape(B, N, C)
    return F.linear(y, p["proj.weight"], p["proj.bias"])


def mlp(x: torch.Tensor, H: int, W: int, p: Params) -> torch.Tensor:
    """cod.py:850-858 with DWConv :1525-1531."""
    B, N, _ = x.shape
    h = F.linear(x, p["fc1.weight"], p["fc1.bias"])
    Ch = h.shape[-1]
    g = h.transpose(1, 2).reshape(B, Ch, H, W)
    g = F.conv2d(g, p["dwconv.dwconv.weight"], p["dwconv.dwconv.bias"], padding=1, groups=Ch)
    h = g.flatten(2).transpose(1, 2)
    h = TDR.gelu_erf(h)
    return F.linear(h, p["fc2.weight"], p["fc2.bias"])


def block(x: torch.Tensor, H: int, W: int, p: Params, heads: int, sr: int) -> torch.Tensor:
    """cod.py:957-961 in eval mode (DropPath = identity)."""
    C = x.shape[-1]
    a = F.layer_norm(x, (C,), p["norm1.weight"], p["norm1.bias"], BLOCK_EPS)
    x = x + attention(a, H, W, sub(p, "attn"), heads, sr)
    a = F.layer_norm(x, (C,), p["norm2.weight"], p["norm2.bias"], BLOCK_EPS)
    return x + mlp(a, H, W, sub(p, "mlp"))


def forward_features(image: torch.Tensor, depth: torch.Tensor, p: Params) -> Tuple[torch.Tensor, List[torch.Tensor]]:
    """cod.py:1455-1509: returns (embedding1, [stage outputs (B, C_s, H_s, W_s)])."""
    emb1, _, tokens = TDR.texture_prompts(image, depth, sub(p, "prompt_encoder"), sub(p, "prompt_decoder"))
    B = image.shape[0]
    outs = []
    x = image
    for s in range(4):
        patch, stride = (7, 4) if s == 0 else (3, 2)
        x, H, W = overlap_patch_embed(x, sub(p, f"patch_embed{s + 1}"), patch, stride)
        for i in range(DEPTHS[s]):
            x = block(x + tokens[s][i].reshape(x.shape), H, W, sub(p, f"block{s + 1}.{i}"), NUM_HEADS[s], SR_RATIOS[s])
        C = x.shape[-1]
        x = F.layer_norm(x, (C,), p[f"norm{s + 1}.weight"], p[f"norm{s + 1}.bias"], BLOCK_EPS)
        x = x.reshape(B, H, W, C).permute(0, 3, 1, 2).contiguous()
        outs.append(x)
    return emb1, outs
