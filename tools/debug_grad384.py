"""Diagnostic: per-parameter gradient error of the fp32 training path at 384^2 (tests/test_gpu_train.py geometry)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import common
from test_gpu_train import _oracle_grads

S, B = int(sys.argv[1]) if len(sys.argv) > 1 else 384, 2
TD = common.package()
enc, dec = TD.build_texture_diffuser(seed=0)
common.perturb_regressor_(enc)
image, depth = common.synthetic_inputs(B, S, seed=3)
grids = common.pvt_token_grids((S, S))
g = torch.Generator().manual_seed(7)
gout_e3 = torch.randn(B, 24, S // 4, S // 4, generator=g) * 1e-2
gout_tok = [[torch.randn(B, grids[s][0] * grids[s][1], e, generator=g) * 1e-2 for _ in range(n)]
            for s, (e, n) in enumerate(zip(common.PVT_EMBED_DIMS, common.PVT_DEPTHS))]
(r1, r3, rtoks), ref = _oracle_grads(enc, dec, image, depth, gout_e3, gout_tok)
enc, dec = enc.cuda().eval(), dec.cuda().eval()
e1, e3, toks = TD.texture_prompts_train(enc, dec, image.cuda(), depth.cuda())
loss = (e3 * gout_e3.cuda()).sum()
for s in range(4):
    for i, t in enumerate(toks[s]):
        loss = loss + (t * gout_tok[s][i].cuda()).sum()
loss.backward()
rows = []
for prefix, mod in (("enc.", enc), ("dec.", dec)):
    for k, p in mod.named_parameters():
        r = ref.get(prefix + k)
        if r is None:
            continue
        rows.append((common.rel_err(p.grad, r), prefix + k, tuple(p.shape), float(r.abs().max())))
rows.sort(reverse=True)
for e, k, sh, mx in rows[:40]:
    print(f"{e:.3e} {k} {sh} max|ref|={mx:.3e}")
print("n>1e-4:", sum(1 for r in rows if r[0] > 1e-4), "of", len(rows))
