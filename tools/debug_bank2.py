"""Diagnostic 3: decoder bank on the REAL emb3 of the 384^2 path: ours vs float64 torch on the same emb3, and the
sensitivity of the float64 gradients to a 1e-6 relative perturbation of emb3 (ReLU-kink conditioning)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, torch.nn.functional as F
import common
TD = common.package()
from dgtd_b200.twig.ops.functions import decoder_bank as DB
from dgtd_b200.twig.ops import capi

def rel(a, b):
    return float((a.detach().double() - b.detach().double()).abs().max() / b.detach().double().abs().max().clamp_min(1e-30))

S, B = 384, 2
enc, dec = TD.build_texture_diffuser(seed=0)
common.perturb_regressor_(enc)
enc, dec = enc.cuda().eval(), dec.cuda().eval()
image, depth = common.synthetic_inputs(B, S, seed=3)
with torch.no_grad():
    _, emb3 = enc._forward_train(image.cuda(), depth.cuda(), capi.F32)
emb3 = emb3.detach().float().contiguous()
print("emb3", tuple(emb3.shape), "strides", emb3.stride(), "absmax", float(emb3.abs().max()), "std", float(emb3.std()))
h = emb3.shape[1]
grids = common.pvt_token_grids((S, S))
g = torch.Generator().manual_seed(7)
cfg = {"stages": [(len(dec[s].decoder), tuple(grids[s])) for s in range(4)], "mode": capi.F32}
params = [t for s in range(4) for d in dec[s].decoder for t in (d.decoder[0].weight, d.decoder[0].bias, d.decoder[2].weight,
                                                                  d.decoder[2].bias, d.decoder[4].weight, d.decoder[4].bias)]
emb = emb3.clone().requires_grad_(True)
outs = DB.DecoderBankFn.apply(emb, cfg, *params)
gouts = [torch.randn(o.shape, generator=g).cuda() * 1e-2 for o in outs]
grads = torch.autograd.grad(sum((o * go).sum() for o, go in zip(outs, gouts)), [emb] + params)

def reference(e64):
    p64 = [p.detach().double().requires_grad_(True) for p in params]
    e64 = e64.requires_grad_(True)
    routs, i = [], 0
    pre_min = []
    for s in range(4):
        for d in range(len(dec[s].decoder)):
            w1, b1, w2, b2, w3, b3 = p64[6 * i:6 * i + 6]
            a1 = F.conv2d(e64, w1, b1, padding=1)
            a2 = F.conv2d(F.relu(a1), w2, b2, padding=1)
            y = F.conv2d(F.relu(a2), w3, b3, padding=1)
            if i == 0:
                pre_min.append((float(a1.abs().median()), float((a1.abs() < 1e-5 * a1.abs().max()).double().mean()),
                                float(a2.abs().median()), float((a2.abs() < 1e-5 * a2.abs().max()).double().mean()), float(a2.abs().max())))
            if tuple(grids[s]) != (h, h):
                y = F.interpolate(y, size=tuple(grids[s]), mode="bilinear")
            routs.append(y.flatten(2).permute(0, 2, 1))
            i += 1
    rg = torch.autograd.grad(sum((o * go.double()).sum() for o, go in zip(routs, gouts)), [e64] + p64)
    return routs, rg, pre_min

e64 = emb3.double().permute(0, 3, 1, 2).contiguous()
routs, rg, pm = reference(e64.clone())
print("decoder 0: |a1| median, frac(|a1| < 1e-5 max), |a2| median, frac(|a2| < 1e-5 max), max|a2|:", pm)
print("fwd", max(rel(a, b) for a, b in zip(outs, routs)), "demb", rel(grads[0].permute(0, 3, 1, 2), rg[0]))
errs = sorted([(rel(a, b), j // 6, j % 6) for j, (a, b) in enumerate(zip(grads[1:], rg[1:]))], reverse=True)
print("ours vs f64 on the same emb3, worst params:", [(f"{e:.1e}", d, k) for e, d, k in errs[:8]])
pert = e64 * (1 + 1e-6 * torch.randn(e64.shape, generator=torch.Generator().manual_seed(1)).double().cuda())
_, rg2, _ = reference(pert)
errs2 = sorted([(rel(a, b), j // 6, j % 6) for j, (a, b) in enumerate(zip(rg2[1:], rg[1:]))], reverse=True)
print("f64 reference under a 1e-6 relative perturbation of emb3: demb", rel(rg2[0], rg[0]), "worst params:",
      [(f"{e:.1e}", d, k) for e, d, k in errs2[:8]])
