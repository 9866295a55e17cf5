"""Diagnostic 2: where does the 384^2 fp32 gradient error enter?  Compares forward emb3 and d loss / d emb3."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import common
from oracle import texture_diffuser_ref as O
from dgtd_b200.twig.ops.functions import train_func as TF

S, B = int(sys.argv[1]) if len(sys.argv) > 1 else 384, int(sys.argv[2]) if len(sys.argv) > 2 else 2
use_tok = (sys.argv[3] if len(sys.argv) > 3 else "all")
TD = common.package()
enc, dec = TD.build_texture_diffuser(seed=0)
common.perturb_regressor_(enc)
image, depth = common.synthetic_inputs(B, S, seed=3)
grids = common.pvt_token_grids((S, S))
g = torch.Generator().manual_seed(7)
gout_e3 = torch.randn(B, 24, S // 4, S // 4, generator=g) * 1e-2
gout_tok = [[torch.randn(B, grids[s][0] * grids[s][1], e, generator=g) * 1e-2 for _ in range(n)]
            for s, (e, n) in enumerate(zip(common.PVT_EMBED_DIMS, common.PVT_DEPTHS))]
pe = {k: v.detach().double().clone().requires_grad_(True) for k, v in enc.state_dict().items()}
pd = {k: v.detach().double().clone().requires_grad_(True) for k, v in dec.state_dict().items()}
e1, e3, toks = O.texture_prompts(image.double(), depth.double(), pe, pd)
e3.retain_grad()
loss = (e3 * gout_e3.double()).sum()
if use_tok == "all":
    for s in range(4):
        for i, t in enumerate(toks[s]):
            loss = loss + (t * gout_tok[s][i].double()).sum()
loss.backward()
ref_de3 = e3.grad
enc, dec = enc.cuda().eval(), dec.cuda().eval()
hook = {}
orig = TF.LayoutFn.apply
e1g, e3g, toksg = TD.texture_prompts_train(enc, dec, image.cuda(), depth.cuda())
e3g.retain_grad()
print("fwd e3 err", common.rel_err(e3g, e3), "tok errs", max(common.rel_err(a, b) for la, lb in zip(toksg, toks) for a, b in zip(la, lb)))
lossg = (e3g * gout_e3.cuda()).sum()
if use_tok == "all":
    for s in range(4):
        for i, t in enumerate(toksg[s]):
            lossg = lossg + (t * gout_tok[s][i].cuda()).sum()
lossg.backward()
rows = []
for prefix, mod, ref in (("enc.", enc, pe), ("dec.", dec, pd)):
    for k, p in mod.named_parameters():
        r = ref[k].grad
        if r is None or p.grad is None:
            continue
        rows.append((common.rel_err(p.grad, r), prefix + k))
rows.sort(reverse=True)
print("tokens in loss:", use_tok, " worst:", rows[:8])
print("n>1e-4:", sum(1 for r in rows if r[0] > 1e-4), "of", len(rows))
# by trunk depth
for name in ["enc.encoder2.convs.0.weight", "enc.encoder2.convs.3.weight", "enc.encoder2.fusion.weight" , "enc.encoder2.stages.3.2.pwconv2.weight",
             "enc.encoder2.stages.3.0.pwconv1.weight", "enc.encoder2.stages.2.26.pwconv2.weight", "enc.encoder2.stages.2.0.pwconv1.weight",
             "enc.encoder2.stages.0.0.pwconv1.weight", "enc.encoder2.downsample_layers.0.0.weight", "enc.encoder1.weight"]:
    for e, k in rows:
        if k == name:
            print(f"  {k}: {e:.2e}")
