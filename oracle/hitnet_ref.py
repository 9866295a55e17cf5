"""CPU restatement of the Hitnet iterative decoder that turns the four backbone maps into the
segmentation logits, and of the `cod` predict head on top of it.

TEST INFRASTRUCTURE ONLY: imported by tests/, tests/golden/make_golden_hitnet.py and bench.py's CPU
legs; never by the product path.  Plain functional torch on a state dict with the reference's keys
(`Hitnet`, cod.py:685-807); pinned against the unmodified reference class by
tests/golden/make_golden_hitnet.py (fixture tests/golden/hitnet_*.npz).  Inference semantics
(BatchNorm uses its running statistics) by default; `train=True` = train-mode BatchNorm (batch
statistics, biased variance; the running buffers are not touched here), pinned against the
unmodified class in train() by the same script (fixture hitnet_train_128.npz: loss + gradients).

Reference lines followed:
  BasicConv2d   cod.py:355-368   conv (no bias) -> BatchNorm2d; the ReLU member is never applied
  CALayer       cod.py:413-429   global mean -> 1x1 (C -> C/r) -> ReLU -> 1x1 (C/r -> C) -> sigmoid -> x * y
  CAB           cod.py:434-451   conv3 -> PReLU -> conv3 -> CALayer -> + x      (no biases; the PReLU
                                 instance is the constructor's default argument, i.e. ONE slope shared by
                                 all CABs of the network, cod.py:686)
  SAM           cod.py:454-506   per input: channel gate fc(mean) and a scalar gate fc_wight(mean);
                                 x_h * y_h * w_h + x_l * y_l * w_l
  Hitnet.forward cod.py:743-807  CIM (2 CABs on x1), Translayers, 4 feedback iterations
                                 (x4 <- compress_out(cat(up4(x4), cfm)), x2 <- compress_out2(cat(x2, cfm))),
                                 out_CFM per iteration x8 bilinear, SAM head x8 bilinear
  resizes       nn.Upsample(align_corners=True) for upsample / upsample_4 / down05 (:709,:733,:737),
                F.interpolate(scale_factor=8, 'bilinear') = align_corners False for the predictions
  cod predict   cod.py:147-149   interpolate(P1[-1] + P2, label size, bilinear, align_corners False)
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import torch
import torch.nn.functional as F

from . import pvt_ref as PVT

Params = Dict[str, torch.Tensor]
BN_EPS = 1e-5


def sub(p: Params, prefix: str) -> Params:
    n = len(prefix) + 1
    return {k[n:]: v for k, v in p.items() if k.startswith(prefix + ".")}


def resize_bilinear(x: torch.Tensor, oh: int, ow: int, align_corners: bool) -> torch.Tensor:
    """Explicit two-tap bilinear resize of an NCHW tensor (no antialiasing), both conventions."""
    def taps(n_in: int, n_out: int):
        d = torch.arange(n_out, dtype=x.dtype)
        if align_corners:
            src = d * ((n_in - 1) / (n_out - 1)) if n_out > 1 else torch.zeros_like(d)
        else:
            src = ((d + 0.5) * (n_in / n_out) - 0.5).clamp_min(0.0)
        i0 = src.floor().long().clamp_max(n_in - 1)
        i1 = (i0 + 1).clamp_max(n_in - 1)
        f = src - i0.to(x.dtype)
        return i0, i1, f
    y0, y1, fy = taps(x.shape[2], oh)
    x0, x1, fx = taps(x.shape[3], ow)
    rows = x[:, :, y0, :] * (1 - fy)[None, None, :, None] + x[:, :, y1, :] * fy[None, None, :, None]
    return rows[:, :, :, x0] * (1 - fx) + rows[:, :, :, x1] * fx


def basic_conv(x: torch.Tensor, p: Params, stride: int = 1, padding: int = 0, train: bool = False) -> torch.Tensor:
    y = F.conv2d(x, p["conv.weight"], None, stride=stride, padding=padding)
    if train:       # nn.BatchNorm2d in train(): statistics of this batch over (B, H, W), biased variance
        mean = y.mean(dim=(0, 2, 3), keepdim=True)
        var = ((y - mean) ** 2).mean(dim=(0, 2, 3), keepdim=True)
        xhat = (y - mean) / torch.sqrt(var + BN_EPS)
        return xhat * p["bn.weight"][None, :, None, None] + p["bn.bias"][None, :, None, None]
    scale = p["bn.weight"] / torch.sqrt(p["bn.running_var"] + BN_EPS)
    shift = p["bn.bias"] - p["bn.running_mean"] * scale
    return y * scale[None, :, None, None] + shift[None, :, None, None]


def channel_gate(mean: torch.Tensor, w1: torch.Tensor, w2: torch.Tensor) -> torch.Tensor:
    """sigmoid(W2 relu(W1 mean)) on (B, C) means; W given as conv (o,i,1,1) or linear (o,i)."""
    w1 = w1.reshape(w1.shape[0], -1)
    w2 = w2.reshape(w2.shape[0], -1)
    return torch.sigmoid(torch.relu(mean @ w1.t()) @ w2.t())


def cab(x: torch.Tensor, p: Params) -> torch.Tensor:
    r = F.conv2d(x, p["body.0.weight"], None, padding=1)
    a = p["body.1.weight"]
    r = torch.where(r >= 0, r, r * a.reshape(1, -1, 1, 1))
    r = F.conv2d(r, p["body.2.weight"], None, padding=1)
    g = channel_gate(r.mean(dim=(2, 3)), p["CA.conv_du.0.weight"], p["CA.conv_du.2.weight"])
    return r * g[:, :, None, None] + x


def cab_pair(x: torch.Tensor, p: Params) -> torch.Tensor:
    return cab(cab(x, sub(p, "0")), sub(p, "1"))


def sam(x_h: torch.Tensor, x_l: torch.Tensor, p: Params) -> torch.Tensor:
    def one(x):
        m = x.mean(dim=(2, 3))
        gate = channel_gate(m, p["fc.0.weight"], p["fc.2.weight"])
        scal = channel_gate(m, p["fc_wight.0.weight"], p["fc_wight.2.weight"])
        return x * gate[:, :, None, None] * scal[:, :, None, None]
    return one(x_h) + one(x_l)


def conv1x1_bias(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    return F.conv2d(x, w, b)


def decode(feats: Sequence[torch.Tensor], p: Params, iterations: int = 4, train: bool = False
           ) -> Tuple[List[torch.Tensor], torch.Tensor]:
    """cod.py:752-805 on the four backbone maps -> ([4 stage predictions], SAM prediction), each (B,1,8h,8w)
    with (h, w) the stride-8 grid."""
    x1, x2, x3, x4 = feats
    cim = cab_pair(x1, sub(p, "decoder_level1"))
    x2_t = basic_conv(x2, sub(p, "Translayer2_1"), train=train)
    x3_t = basic_conv(x3, sub(p, "Translayer3_1"), train=train)
    x4_t = basic_conv(x4, sub(p, "Translayer4_1"), train=train)
    preds: List[torch.Tensor] = []
    cfm = None
    for it in range(iterations):
        if cfm is not None:
            up = resize_bilinear(x4_t, 4 * x4_t.shape[2], 4 * x4_t.shape[3], True)
            x4_t = basic_conv(torch.cat((up, cfm), 1), sub(p, "compress_out"), stride=4, padding=2, train=train)
        x4_f = cab_pair(x4_t, sub(p, "decoder_level4"))
        up = resize_bilinear(x4_f, 2 * x4_f.shape[2], 2 * x4_f.shape[3], True)
        x3_f = cab_pair(torch.cat((x3_t, up), 1), sub(p, "decoder_level3"))
        if it > 0:
            x2_t = basic_conv(torch.cat((x2_t, cfm), 1), sub(p, "compress_out2"), train=train)
        up = resize_bilinear(x3_f, 2 * x3_f.shape[2], 2 * x3_f.shape[3], True)
        x2_f = cab_pair(torch.cat((x2_t, up), 1), sub(p, "decoder_level2"))
        cfm = basic_conv(x2_f, sub(p, "conv4"), padding=1, train=train)
        pr = conv1x1_bias(cfm, p["out_CFM.weight"], p["out_CFM.bias"])
        preds.append(resize_bilinear(pr, 8 * pr.shape[2], 8 * pr.shape[3], False))
    t2 = basic_conv(cim, sub(p, "Translayer2_0"), train=train)
    t2 = resize_bilinear(t2, t2.shape[2] // 2, t2.shape[3] // 2, True)
    s = sam(cfm, t2, sub(p, "SAM"))
    pr = conv1x1_bias(s, p["out_SAM.weight"], p["out_SAM.bias"])
    return preds, resize_bilinear(pr, 8 * pr.shape[2], 8 * pr.shape[3], False)


def hitnet_forward(image: torch.Tensor, depth: torch.Tensor, p: Params, train: bool = False):
    """`Hitnet.forward(x, pred_normal)` (cod.py:743-807) -> (embedding1, [P1 x4], P2).  `train`: the decoder's
    BatchNorms use batch statistics (the backbone has no BatchNorm; its DropPath is the identity here)."""
    emb1, feats = PVT.forward_features(image, depth, sub(p, "backbone"))
    preds, p2 = decode(feats, p, train=train)
    return emb1, preds, p2


def predict_logits(preds: Sequence[torch.Tensor], p2: torch.Tensor, size: Sequence[int]) -> torch.Tensor:
    """cod.py:149: the map whose sigmoid is thresholded / scored."""
    out = preds[-1] + p2
    return resize_bilinear(out, int(size[0]), int(size[1]), False)
