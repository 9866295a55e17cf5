"""Diagnostic: DecoderBankFn (fp32 mode) input / parameter gradients vs float64 torch autograd on the GPU, by size."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, torch.nn.functional as F
import common
TD = common.package()
from dgtd_b200.twig.ops.functions import decoder_bank as DB
from dgtd_b200.twig.ops import capi

def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))

for h in (24, 48, 96):
    torch.manual_seed(0)
    enc, dec = TD.build_texture_diffuser(seed=0)
    dec = dec.cuda()
    B = 2
    g = torch.Generator().manual_seed(h)
    emb = torch.randn(B, h, h, 24, generator=g).cuda().requires_grad_(True)
    grids = [(h, h), (h // 2, h // 2), (h // 4, h // 4), (h // 8, h // 8)]
    cfg = {"stages": [(len(dec[s].decoder), grids[s]) for s in range(4)], "mode": capi.F32}
    params = [t for s in range(4) for d in dec[s].decoder for t in (d.decoder[0].weight, d.decoder[0].bias, d.decoder[2].weight,
                                                                      d.decoder[2].bias, d.decoder[4].weight, d.decoder[4].bias)]
    outs = DB.DecoderBankFn.apply(emb, cfg, *params)
    gouts = [torch.randn(o.shape, generator=g).cuda() * 1e-2 for o in outs]
    loss = sum((o * go).sum() for o, go in zip(outs, gouts))
    grads = torch.autograd.grad(loss, [emb] + params)
    # float64 reference
    emb64 = emb.detach().double().permute(0, 3, 1, 2).requires_grad_(True)
    p64 = [p.detach().double().requires_grad_(True) for p in params]
    routs, i = [], 0
    for s in range(4):
        for d in range(len(dec[s].decoder)):
            w1, b1, w2, b2, w3, b3 = p64[6 * i:6 * i + 6]
            y = F.conv2d(F.relu(F.conv2d(F.relu(F.conv2d(emb64, w1, b1, padding=1)), w2, b2, padding=1)), w3, b3, padding=1)
            if grids[s] != (h, h):
                y = F.interpolate(y, size=grids[s], mode="bilinear")
            routs.append(y.flatten(2).permute(0, 2, 1))
            i += 1
    fwd = max(rel(a, b) for a, b in zip(outs, routs))
    rloss = sum((o * go.double()).sum() for o, go in zip(routs, gouts))
    rg = torch.autograd.grad(rloss, [emb64] + p64)
    e_emb = rel(grads[0].permute(0, 3, 1, 2), rg[0])
    errs = [(rel(a, b), j) for j, (a, b) in enumerate(zip(grads[1:], rg[1:]))]
    errs.sort(reverse=True)
    print(f"h={h}: fwd {fwd:.2e} demb {e_emb:.2e} worst params {[(f'{e:.1e}', j // 6, j % 6) for e, j in errs[:6]]}")
