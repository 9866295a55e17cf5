"""structure loss + deep supervision on the CUDA path vs the reference golden / the oracle."""
import os

import numpy as np
import pytest
import torch

import common
from oracle import loss_ref as L

pytestmark = pytest.mark.gpu


def losses():
    common.package()
    from dgtd_b200.twig.model import losses as M
    return M


@pytest.mark.parametrize("tag,shape", [("a", (2, 48, 64)), ("b", (3, 40, 40))])
def test_cal_loss_value_and_gradient_match_reference_golden(tag, shape):
    M = losses()
    g = np.load(os.path.join(common.GOLDEN, "loss_small.npz"))
    preds, gts = common.loss_inputs(*shape, seed=ord(tag))
    p = preds.cuda().requires_grad_(True)
    loss = M.cal_loss(p, gts.cuda())
    loss.backward()
    assert abs(float(loss.detach()) - float(g[f"{tag}_loss"])) <= 1e-5 * abs(float(g[f"{tag}_loss"]))
    ref = torch.from_numpy(g[f"{tag}_grad"])
    err = float((p.grad.cpu().double() - ref).abs().max() / ref.abs().max())
    assert err <= 1e-4, err


def test_boundary_weight_and_full_size_against_oracle():
    M = losses()
    preds, gts = common.loss_inputs(2, 384, 384, seed=5)
    w = M.boundary_weight(gts.cuda())
    wr = L.boundary_weight(gts.double())
    assert float((w.cpu().double() - wr).abs().max()) <= 1e-5
    p = preds.cuda().requires_grad_(True)
    loss = M.cal_loss(p, gts.cuda())
    loss.backward()
    q = preds.double().requires_grad_(True)
    ref = L.structure_loss(q, gts.double())
    (gr,) = torch.autograd.grad(ref, q)
    assert abs(float(loss.detach()) - float(ref.detach())) <= 1e-5 * abs(float(ref.detach()))
    assert float((p.grad.cpu().double() - gr).abs().max() / gr.abs().max()) <= 1e-4
    loss2 = M.cal_loss(p.detach(), gts.cuda())
    assert float(loss2.detach()) == float(loss.detach())            # fixed-order reductions: bit-stable


def test_deep_supervision_sum():
    M = losses()
    g = torch.Generator().manual_seed(9)
    _, gts = common.loss_inputs(2, 96, 96, seed=3)
    P1 = [2.0 * torch.randn(2, 1, 96, 96, generator=g) for _ in range(4)]
    P2 = 2.0 * torch.randn(2, 1, 96, 96, generator=g)
    leaves = [t.cuda().requires_grad_(True) for t in P1 + [P2]]
    loss = M.deep_supervision_loss(leaves[:4], leaves[4], gts.cuda())
    loss.backward()
    ql = [t.double().requires_grad_(True) for t in P1 + [P2]]
    ref = L.deep_supervision_loss(ql[:4], ql[4], gts.double())
    gr = torch.autograd.grad(ref, ql, allow_unused=True)
    assert abs(float(loss.detach()) - float(ref.detach())) <= 1e-5 * abs(float(ref.detach()))
    assert leaves[0].grad is None and gr[0] is None or float(gr[0].abs().max()) == 0.0   # it = 0 has weight 0
    for a, b in zip(leaves[1:], gr[1:]):
        assert float((a.grad.cpu().double() - b).abs().max() / b.abs().max()) <= 1e-4


def test_ssim_constant_matches_the_reference_golden_and_the_oracle():
    """cod.py:142-144 / SSIM :316-351: value recorded from the unmodified reference, plus a full-size case vs the
    oracle (batch-wide min-max, reflection padding at every border)."""
    import os
    import numpy as np
    from oracle import loss_ref as L
    common.package()
    from dgtd_b200.twig.model import losses as M
    g = np.load(os.path.join(common.GOLDEN, "loss_small.npz"))
    emb, img = torch.from_numpy(g["ssim_emb"]), torch.from_numpy(g["ssim_img"])
    got = float(M.ssim_constant(emb.float().cuda(), img.float().cuda()))
    assert abs(got - float(g["ssim_value"])) <= 1e-5 * float(g["ssim_value"])
    gen = torch.Generator().manual_seed(9)
    e = torch.rand(3, 3, 384, 352, generator=gen) * 2 - 0.3
    y = torch.randn(3, 3, 384, 352, generator=gen)
    got = float(M.ssim_constant(e.cuda(), y.cuda()))
    want = float(L.ssim_constant(e.double(), y.double()))
    assert abs(got - want) <= 1e-5 * want
    a = M.ssim_constant(e.cuda(), y.cuda())
    assert float(a) == got                      # fixed-order reductions: bit-stable
