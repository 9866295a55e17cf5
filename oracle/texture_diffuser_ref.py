"""CPU oracle for the depth-guided texture-diffusion hot path.

TEST INFRASTRUCTURE ONLY -- imported by `tests/`, `__graft_entry__.smoke()` and by the
`cpu_baseline` / `--impl reference` legs of `bench.py`, never by the product package.

This is a functional restatement (torch on CPU, any float dtype; float64 is the master
precision) of the reference algorithm in `/root/reference/twig/model/cod.py`.  Every
function cites the reference lines it follows.  Parameters are passed as a flat
``dict[str, Tensor]`` whose keys are the reference ``state_dict`` keys relative to the
module (e.g. ``"encoder2.stages.2.5.pwconv1.weight"``), so the same dict drives the
reference module, the oracle and the CUDA path.

Pinning: the reference ships no golden vectors or tests for this path, so
the oracle is pinned against outputs of the reference module itself executed in the
authoring container (`tests/golden/make_golden.py` -> `tests/golden/*.npz`, checked by
`tests/test_oracle_golden.py`).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Params = Dict[str, torch.Tensor]

CONVNEXT_DIMS = (128, 256, 512, 1024)      # cod.py:1125
CONVNEXT_DEPTHS = (3, 3, 27, 3)            # cod.py:1141
LATENT_DIM = 24                            # cod.py:1394
GRID = 12                                  # cod.py:1283
KSIZE = 7                                  # cod.py:1181
MAX_STEP = 4                               # cod.py:1181
FREQ_RATE = 0.3                            # cod.py:1254
PVT_EMBED_DIMS = (64, 128, 320, 512)       # cod.py:1785 (pvt_v2_b2)
PVT_DEPTHS = (3, 4, 6, 3)                  # cod.py:1786


def sub(params: Params, prefix: str) -> Params:
    """Sub-dict of `params` under `prefix.` with the prefix stripped."""
    p = prefix + "."
    return {k[len(p):]: v for k, v in params.items() if k.startswith(p)}


# --------------------------------------------------------------------------- a1
def surface_normals(depth: torch.Tensor) -> torch.Tensor:
    """cod.py:96-109.  Central differences inside, one-sided at the borders
    (`torch.gradient`, unit spacing), normal = (-dz/d(dim2), -dz/d(dim3), 1)/norm."""
    d = depth
    gh = torch.empty_like(d)
    gh[:, :, 1:-1] = (d[:, :, 2:] - d[:, :, :-2]) * 0.5
    gh[:, :, 0] = d[:, :, 1] - d[:, :, 0]
    gh[:, :, -1] = d[:, :, -1] - d[:, :, -2]
    gw = torch.empty_like(d)
    gw[..., 1:-1] = (d[..., 2:] - d[..., :-2]) * 0.5
    gw[..., 0] = d[..., 1] - d[..., 0]
    gw[..., -1] = d[..., -1] - d[..., -2]
    nx, ny, nz = -gh, -gw, torch.ones_like(d)
    norm = torch.sqrt(nx * nx + ny * ny + nz * nz)
    return torch.cat((nx / norm, ny / norm, nz / norm), dim=1)


# --------------------------------------------------------------------------- a2
def highpass_line(h: int, w: int, rate: float = FREQ_RATE) -> int:
    """cod.py:1261 (`line = int((w*h*rate)**.5 // 2)`)."""
    return int((w * h * rate) ** 0.5 // 2)


def fft_highpass(x: torch.Tensor, rate: float = FREQ_RATE) -> torch.Tensor:
    """cod.py:1256-1271.  Zero the centred (2*line)^2 block of the shifted spectrum and
    return |Re(ifft2)|.  `fftshift` with no `dim` shifts *all* dims; the matching
    `ifftshift` undoes the batch/channel roll, so only the spatial shift matters."""
    h, w = x.shape[-2:]
    line = highpass_line(h, w, rate)
    spec = torch.fft.fft2(x, norm="forward")
    spec = torch.fft.fftshift(spec, dim=(-2, -1))
    keep = torch.ones(h, w, dtype=x.dtype)
    keep[h // 2 - line:h // 2 + line, w // 2 - line:w // 2 + line] = 0
    spec = spec * keep
    spec = torch.fft.ifftshift(spec, dim=(-2, -1))
    return torch.fft.ifft2(spec, norm="forward").real.abs()


def lowpass_projector(n: int, line: int, dtype=torch.float64) -> Tuple[torch.Tensor, torch.Tensor]:
    """Real and imaginary part of P = E E^H / n, E[m,k] = exp(2 pi i m k / n) for the
    removed frequency band k in [-line, line-1] (cod.py:1262: rows n/2-line .. n/2+line-1
    of the shifted spectrum).  Re P is a symmetric circulant; Im P comes only from the
    unpaired bin k=-line:  Im P[a,b] = -sin(2 pi line (a-b)/n)/n."""
    d = torch.arange(n, dtype=torch.float64)
    diff = d[:, None] - d[None, :]
    ks = torch.arange(-line, line, dtype=torch.float64)
    ang = 2.0 * math.pi * diff[..., None] * ks / n
    re = torch.cos(ang).sum(-1) / n
    im = torch.sin(ang).sum(-1) / n
    return re.to(dtype), im.to(dtype)


def fft_highpass_projector(x: torch.Tensor, rate: float = FREQ_RATE) -> torch.Tensor:
    """Same operator as `fft_highpass`, written the way the CUDA path computes it:
    |x - (Re P_h x Re P_w^T - Im P_h x Im P_w^T)|  (SURVEY.md appendix A)."""
    h, w = x.shape[-2:]
    line = highpass_line(h, w, rate)
    ah, bh = lowpass_projector(h, line, x.dtype)
    aw, bw = lowpass_projector(w, line, x.dtype)
    low = ah @ x @ aw.transpose(0, 1) - bh @ x @ bw.transpose(0, 1)
    return (x - low).abs()


# --------------------------------------------------------------------------- a3..a6
def nearest_grid(x: torch.Tensor, grid: int = GRID) -> torch.Tensor:
    """cod.py:1295, `F.interpolate(x, size=[12,12])` (nearest): src = floor(dst*in/out)."""
    h, w = x.shape[-2:]
    iy = torch.div(torch.arange(grid) * h, grid, rounding_mode="floor")
    ix = torch.div(torch.arange(grid) * w, grid, rounding_mode="floor")
    return x[:, :, iy][:, :, :, ix]


def regress_weights(xx: torch.Tensor, reg_w: torch.Tensor, reg_b: torch.Tensor) -> torch.Tensor:
    """cod.py:1058-1060: sigmoid(1x1 conv 3 -> latent*49); channel = c*49 + ki*7 + kj."""
    return torch.sigmoid(F.conv2d(xx, reg_w, reg_b))


def bilinear_resize(x: torch.Tensor, size: Sequence[int]) -> torch.Tensor:
    """`F.interpolate(mode='bilinear')`, align_corners=False, no antialias, restated
    explicitly: src = (dst+0.5)*in/out - 0.5 clamped at 0, neighbour index clamped."""
    n, c, h, w = x.shape
    oh, ow = int(size[0]), int(size[1])

    def axis(inp: int, out: int):
        s = (torch.arange(out, dtype=torch.float64) + 0.5) * (inp / out) - 0.5
        s = s.clamp_min(0.0)
        i0 = s.floor().to(torch.int64).clamp_max(inp - 1)
        i1 = (i0 + 1).clamp_max(inp - 1)
        l1 = (s - i0.to(torch.float64)).to(x.dtype)
        return i0, i1, l1

    y0, y1, ly = axis(h, oh)
    x0, x1, lx = axis(w, ow)
    rows = x[:, :, y0] * (1 - ly)[None, None, :, None] + x[:, :, y1] * ly[None, None, :, None]
    return rows[..., x0] * (1 - lx) + rows[..., x1] * lx


def depth_to_grid(depth: torch.Tensor, enc_w: torch.Tensor, enc_b: torch.Tensor,
                  grid: int = GRID) -> torch.Tensor:
    """cod.py:1297-1298: `encoder1` (1x1 conv 1 -> latent) at full resolution, then
    bilinear down-sample to grid x grid."""
    return bilinear_resize(F.conv2d(depth, enc_w, enc_b), (grid, grid))


def message_passing_core(x: torch.Tensor, weight: torch.Tensor, k: int = KSIZE,
                         max_step: int = MAX_STEP, eps: float = 1e-5) -> torch.Tensor:
    """cod.py:1190-1205 restated as an explicit stencil (no unfold).

    x: (n,c,h,w); weight: (n, wc*k*k, h, w) with wc in {1, c}; channel = cw*k*k + ki*k + kj.
    Wn = W / (sum_k W + eps); `max_step` times x[c,p] <- sum_k Wn[c,k,p] * x[c, p+delta_k]
    with zero padding and no re-normalisation at the borders."""
    n, c, h, w = x.shape
    kk = k * k
    wc = weight.shape[1] // kk
    wt = weight.reshape(n, wc, kk, h, w)
    wn = wt / (wt.sum(2, keepdim=True) + eps)
    r = k // 2
    steps = max(h, w) if max_step < 0 else max_step
    for _ in range(steps):
        xp = F.pad(x, (r, r, r, r))
        acc = torch.zeros_like(x)
        for ki in range(k):
            for kj in range(k):
                acc = acc + wn[:, :, ki * k + kj] * xp[:, :, ki:ki + h, kj:kj + w]
        x = acc
    return x


def message_passing(x: torch.Tensor, weight: torch.Tensor, conv_w: torch.Tensor,
                    conv_b: torch.Tensor, img_size: Sequence[int], k: int = KSIZE,
                    max_step: int = MAX_STEP) -> torch.Tensor:
    """cod.py:1189-1208: diffusion core, 1x1 conv latent -> 3, bilinear up to img_size."""
    y = message_passing_core(x, weight, k, max_step)
    y = F.conv2d(y, conv_w, conv_b)
    return bilinear_resize(y, img_size)


# --------------------------------------------------------------------------- a12
def layer_norm_channels_first(x, w, b, eps: float = 1e-6):
    """cod.py:1044-1049 (biased variance, eps inside the sqrt)."""
    u = x.mean(1, keepdim=True)
    s = (x - u).pow(2).mean(1, keepdim=True)
    return w[:, None, None] * ((x - u) / torch.sqrt(s + eps)) + b[:, None, None]


def layer_norm_channels_last(x, w, b, eps: float = 1e-6):
    """cod.py:1042-1043 (`F.layer_norm` over the trailing channel dim)."""
    u = x.mean(-1, keepdim=True)
    s = (x - u).pow(2).mean(-1, keepdim=True)
    return (x - u) / torch.sqrt(s + eps) * w + b


def gelu_erf(x):
    """`nn.GELU()` default (exact erf form), cod.py:1098."""
    return 0.5 * x * (1.0 + torch.erf(x * (1.0 / math.sqrt(2.0))))


# --------------------------------------------------------------------------- a8
def convnext_block(x: torch.Tensor, p: Params, keep_scale: Optional[torch.Tensor] = None) -> torch.Tensor:
    """cod.py:1104-1117.  `keep_scale` is the per-sample DropPath factor (mask/keep_prob,
    shape (B,)); None == eval mode."""
    c = x.shape[1]
    y = F.conv2d(x, p["dwconv.weight"], p["dwconv.bias"], padding=3, groups=c)
    y = y.permute(0, 2, 3, 1)
    y = layer_norm_channels_last(y, p["norm.weight"], p["norm.bias"])
    y = F.linear(y, p["pwconv1.weight"], p["pwconv1.bias"])
    y = gelu_erf(y)
    y = F.linear(y, p["pwconv2.weight"], p["pwconv2.bias"])
    if "gamma" in p:
        y = p["gamma"] * y
    y = y.permute(0, 3, 1, 2)
    if keep_scale is not None:
        y = y * keep_scale.reshape(-1, 1, 1, 1)
    return x + y


# --------------------------------------------------------------------------- a7, a9
def shape_prop_encoder_pyramid(x: torch.Tensor, p: Params,
                               keep_scales: Optional[List[torch.Tensor]] = None) -> List[torch.Tensor]:
    """cod.py:1165-1169: stem / downsample + ConvNeXt stages; returns the 4 pyramid maps."""
    outs = []
    blk = 0
    for i in range(4):
        d = sub(p, f"downsample_layers.{i}")
        if i == 0:
            x = F.conv2d(x, d["0.weight"], d["0.bias"], stride=4)
            x = layer_norm_channels_first(x, d["1.weight"], d["1.bias"])
        else:
            x = layer_norm_channels_first(x, d["0.weight"], d["0.bias"])
            x = F.conv2d(x, d["1.weight"], d["1.bias"], stride=2)
        for j in range(CONVNEXT_DEPTHS[i]):
            ks = None if keep_scales is None else keep_scales[blk]
            x = convnext_block(x, sub(p, f"stages.{i}.{j}"), ks)
            blk += 1
        outs.append(x)
    return outs


def shape_prop_encoder_head(outs: List[torch.Tensor], p: Params) -> torch.Tensor:
    """cod.py:1171-1177: per-level 1x1 conv -> bilinear to level-0 size -> cat -> 1x1 conv."""
    size = outs[0].shape[2:]
    tmp = [bilinear_resize(F.conv2d(o, p[f"convs.{i}.weight"], p[f"convs.{i}.bias"]), size)
           for i, o in enumerate(outs)]
    return F.conv2d(torch.cat(tmp, 1), p["fusion_conv.weight"], p["fusion_conv.bias"])


def shape_prop_encoder(x: torch.Tensor, p: Params,
                       keep_scales: Optional[List[torch.Tensor]] = None) -> torch.Tensor:
    """cod.py:1163-1177."""
    return shape_prop_encoder_head(shape_prop_encoder_pyramid(x, p, keep_scales), p)


# --------------------------------------------------------------------------- a10
def shape_prop_decoder(emb: torch.Tensor, p: Params) -> torch.Tensor:
    """cod.py:1216-1226: conv3x3+ReLU, conv3x3+ReLU, conv3x3 (pad 1)."""
    y = F.relu(F.conv2d(emb, p["decoder.0.weight"], p["decoder.0.bias"], padding=1))
    y = F.relu(F.conv2d(y, p["decoder.2.weight"], p["decoder.2.bias"], padding=1))
    return F.conv2d(y, p["decoder.4.weight"], p["decoder.4.bias"], padding=1)


def prompt_decoder(emb: torch.Tensor, p: Params, depth: int) -> List[torch.Tensor]:
    """cod.py:1316-1323: `depth` independent decoders on the same embedding."""
    return [shape_prop_decoder(emb, sub(p, f"decoder.{i}")) for i in range(depth)]


# --------------------------------------------------------------------------- prompt_encoder
def prompt_encoder(image: torch.Tensor, cues: torch.Tensor, p: Params,
                   keep_scales: Optional[List[torch.Tensor]] = None,
                   return_intermediates: bool = False):
    """cod.py:1281-1306.  Returns (embedding1 = fft high-pass of the image, embedding3).
    Up-sampling target is the image size (the reference hard-codes 384, cod.py:1252)."""
    x = fft_highpass(image, FREQ_RATE)                                        # :1288
    xx = nearest_grid(x, GRID)                                                # :1295
    weights = regress_weights(xx, p["propagation_weight_regressor.reg.weight"],
                              p["propagation_weight_regressor.reg.bias"])      # :1296
    d12 = depth_to_grid(cues, p["encoder1.weight"], p["encoder1.bias"], GRID)  # :1297-1298
    emb2 = message_passing(d12, weights, p["message_passing.conv.weight"],
                           p["message_passing.conv.bias"], image.shape[-2:])   # :1298
    emb3 = shape_prop_encoder(emb2 + image, sub(p, "encoder2"), keep_scales)   # :1302
    if return_intermediates:
        return x, emb3, dict(xx=xx, weights=weights, d12=d12, emb2=emb2)
    return x, emb3


# --------------------------------------------------------------------------- a11
def pvt_token_grids(img_hw: Sequence[int]) -> List[Tuple[int, int]]:
    """Token grid (H_s, W_s) of the four PVT-v2 stages (OverlapPatchEmbed: 7/4/3, then
    3/2/1 three times; cod.py:1350-1357, conv output size floor((n+2p-k)/s)+1)."""
    h, w = int(img_hw[0]), int(img_hw[1])
    out = []
    h, w = (h + 6 - 7) // 4 + 1, (w + 6 - 7) // 4 + 1
    out.append((h, w))
    for _ in range(3):
        h, w = (h + 2 - 3) // 2 + 1, (w + 2 - 3) // 2 + 1
        out.append((h, w))
    return out


def prompt_to_tokens(prompt: torch.Tensor, hw: Sequence[int]) -> torch.Tensor:
    """cod.py:1471: bilinear resize to the stage grid, then (B,E,H,W) -> (B,H*W,E)."""
    y = bilinear_resize(prompt, hw)
    return y.flatten(2).permute(0, 2, 1).contiguous()


def texture_prompts(image: torch.Tensor, cues: torch.Tensor, enc: Params,
                    dec: Params) -> Tuple[torch.Tensor, torch.Tensor, List[List[torch.Tensor]]]:
    """The hot path as `forward_features` drives it (cod.py:1467-1505) minus the PVT
    blocks: returns (embedding1, embedding3, tokens[s][i]) where tokens[s][i] is the
    (B, H_s*W_s, E_s) tensor that is added to the stage-s token stream before block i.
    `dec` holds the `prompt_decoder` Sequential's keys (`{s}.decoder.{i}.decoder.{0,2,4}.*`)."""
    emb1, emb3 = prompt_encoder(image, cues, enc)
    grids = pvt_token_grids(image.shape[-2:])
    tokens = []
    for s in range(4):
        ps = prompt_decoder(emb3, sub(dec, str(s)), PVT_DEPTHS[s])
        tokens.append([prompt_to_tokens(q, grids[s]) for q in ps])
    return emb1, emb3, tokens
