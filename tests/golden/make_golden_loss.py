"""Golden fixture of the structure loss from the UNMODIFIED reference `cod.cal_loss`
(cod.py:75-84), float64 on CPU, incl. its autograd gradient.  python tests/golden/make_golden_loss.py"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle.ref_loader import load_reference  # noqa: E402
from oracle import loss_ref as L  # noqa: E402
import common  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    m = load_reference()
    rec = {}
    for tag, (B, H, W) in {"a": (2, 48, 64), "b": (3, 40, 40)}.items():
        preds, gts = common.loss_inputs(B, H, W, seed=ord(tag))
        p = preds.double().requires_grad_(True)
        ref = m.cod.cal_loss(None, p, gts.double())
        (g,) = torch.autograd.grad(ref, p)
        q = preds.double().requires_grad_(True)
        mine = L.structure_loss(q, gts.double())
        (g2,) = torch.autograd.grad(mine, q)
        assert abs(float((mine - ref).detach())) < 1e-12 and float((g - g2).abs().max()) < 1e-14
        rec[f"{tag}_loss"] = np.array(float(ref.detach()))
        rec[f"{tag}_grad"] = g.numpy()
    # SSIM constant (cod.py:142-144, SSIM :316-351) and the full `cod.forward(mode='loss')` combination
    g = torch.Generator().manual_seed(5)
    emb = torch.rand(2, 3, 40, 56, generator=g, dtype=torch.float64) * 0.7
    img = torch.randn(2, 3, 40, 56, generator=g, dtype=torch.float64)
    ssim = m.SSIM()
    e_n = (emb - emb.min()) / (emb.max() - emb.min() + 1e-8)
    ref = ssim(e_n, img)
    assert abs(float(ref - L.ssim_constant(emb, img))) < 1e-13
    rec["ssim_emb"], rec["ssim_img"], rec["ssim_value"] = emb.numpy(), img.numpy(), np.array(float(ref))
    np.savez_compressed(os.path.join(OUT, "loss_small.npz"), **rec)
    print({k: (v.shape if v.ndim else float(v)) for k, v in rec.items()})


if __name__ == "__main__":
    main()
