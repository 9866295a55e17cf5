"""Seeded inputs / parameter recipes shared by the tests, the golden generator and bench.py."""
from __future__ import annotations

import json
import os
import sys
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")

PVT_EMBED_DIMS = (64, 128, 320, 512)
PVT_DEPTHS = (3, 4, 6, 3)


def synthetic_inputs(B: int, S: int, seed: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
    """ImageNet-normalised image ~ N(0,1), depth ~ U[0,1) ('L' PNG via ToTensor)."""
    g = torch.Generator("cpu").manual_seed(seed)
    image = torch.randn(B, 3, S, S, generator=g)
    depth = torch.rand(B, 1, S, S, generator=g)
    return image, depth


def perturb_regressor_(pe) -> None:
    """Non-trivial diffusion weights: at random init the regressor logits are ~0 (sigma ~ 0.5
    everywhere) and a box-filter bug would pass; scale x20 and randomise the bias."""
    g = torch.Generator("cpu").manual_seed(1)
    reg = pe.propagation_weight_regressor.reg
    with torch.no_grad():
        reg.weight.mul_(20.0)
        reg.bias.copy_(torch.randn(reg.bias.shape, generator=g))


def pvt_token_grids(img_hw: Sequence[int]) -> List[Tuple[int, int]]:
    h, w = int(img_hw[0]), int(img_hw[1])
    h, w = (h + 6 - 7) // 4 + 1, (w + 6 - 7) // 4 + 1
    out = [(h, w)]
    for _ in range(3):
        h, w = (h + 2 - 3) // 2 + 1, (w + 2 - 3) // 2 + 1
        out.append((h, w))
    return out


def flatten_outputs(e1, e3, toks) -> Dict[str, torch.Tensor]:
    out = {"embedding1": e1, "embedding3": e3}
    for s, lst in enumerate(toks):
        for i, t in enumerate(lst):
            out[f"tokens.{s}.{i}"] = t
    return out


def subsample(key: str, t: torch.Tensor) -> torch.Tensor:
    """Deterministic sub-sample small enough to commit (a few thousand values per tensor)."""
    if key == "embedding1":
        return t[:, :, ::8, ::8]
    if key == "embedding3":
        return t[:, :, ::4, ::4]
    n = t.shape[1]                      # tokens (B, HW, E)
    step = max(1, n // 64)
    return t[:, ::step, ::4]


def moments(t: torch.Tensor) -> np.ndarray:
    t = t.detach().double()
    return np.array([float(t.mean()), float(t.abs().mean()), float((t * t).mean())])


def load_params_fixture() -> dict:
    with open(os.path.join(GOLDEN, "params_seed0.json")) as f:
        return json.load(f)


def package():
    import dgtd_b200  # noqa: F401  (root-level alias module)
    from dgtd_b200.twig.model import texture_diffuser
    return texture_diffuser


def oracle_params(enc, dec) -> Tuple[dict, dict]:
    """state_dicts of the product modules as float64 CPU dicts for the oracle."""
    e = {k: v.detach().double().cpu() for k, v in enc.state_dict().items()}
    d = {k: v.detach().double().cpu() for k, v in dec.state_dict().items()}
    return e, d


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max|a-b| / max|b| (the parity metric of SURVEY.md 8c)."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def fill_params_(module, seed: int = 0) -> None:
    """Constructor-independent parameter recipe (used for the PVT backbone fixtures): every tensor of the
    state dict is drawn from its own generator seeded by (seed, crc32(name)), so the reference classes and
    the mirror classes get identical values whatever their constructors did to the global RNG.
    LayerNorm-like weights ~ 1 + 0.1 N(0,1), biases ~ 0.1 N(0,1) (non-zero on purpose), matrices / conv
    kernels ~ N(0, 1/sqrt(fan_in))."""
    import zlib
    with torch.no_grad():
        for name, t in sorted(module.state_dict().items()):
            if not t.dtype.is_floating_point:
                continue
            g = torch.Generator("cpu").manual_seed((seed * 1000003 + zlib.crc32(name.encode())) % (2 ** 31))
            r = torch.randn(t.shape, generator=g, dtype=torch.float32)
            if name.endswith("running_var"):
                v = (1.0 + 0.3 * r).abs() + 0.05          # BatchNorm variance: positive, not all ~1
            elif t.dim() == 1 and t.numel() == 1:
                v = torch.full(t.shape, 0.2)               # the shared nn.PReLU() slope (cod.py:686)
            elif t.dim() == 1:
                is_scale = name.endswith("weight") or name.endswith("gamma")
                v = 1.0 + 0.1 * r if is_scale else 0.1 * r
            else:
                fan_in = t[0].numel()
                v = r / max(1.0, fan_in) ** 0.5
            t.copy_(v.to(t.dtype))


def hitnet_fixture_params_(net, seed: int = 0) -> None:
    """`fill_params_` + a shift of the CFM head bias so that the predict logits straddle zero (with the plain
    recipe every logit of the 128^2 fixture is negative and the binarised masks would be trivially empty)."""
    fill_params_(net, seed=seed)
    with torch.no_grad():
        net.out_CFM.bias.add_(3.3)


def loss_inputs(B: int, H: int, W: int, seed: int = 0):
    """Logits ~ 3 N(0,1) and a blobby binary-ish ground truth (rectangles + soft edge) for the structure loss."""
    g = torch.Generator("cpu").manual_seed(seed)
    preds = 3.0 * torch.randn(B, 1, H, W, generator=g)
    gts = torch.zeros(B, 1, H, W)
    for b in range(B):
        for _ in range(3):
            y0, x0 = int(torch.randint(0, H - 8, (1,), generator=g)), int(torch.randint(0, W - 8, (1,), generator=g))
            hh, ww = int(torch.randint(4, H // 2, (1,), generator=g)), int(torch.randint(4, W // 2, (1,), generator=g))
            gts[b, 0, y0:y0 + hh, x0:x0 + ww] = 1.0
    gts = (gts + 0.1 * torch.rand(B, 1, H, W, generator=g)).clamp(0, 1)
    return preds, gts
