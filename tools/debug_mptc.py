import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import common
from oracle import texture_diffuser_ref as O
common.package()
from dgtd_b200.twig.ops.functions import texture_diffusion_func as OP
impl = sys.argv[1]
n, h, w, c, T = 1, int(sys.argv[2]), int(sys.argv[3]), 256, 1
g = torch.Generator().manual_seed(23)
x = torch.randn(n, c, h, w, generator=g).to(torch.bfloat16)
wgt = torch.rand(n, 49, h, w, generator=g)
ref = O.message_passing_core(x.double(), wgt.double(), 7, T)
got = OP.message_passing_tiled(x.permute(0, 2, 3, 1).contiguous().cuda(), wgt.cuda(), T, impl=impl)
torch.cuda.synchronize()
got = got.float().permute(0, 3, 1, 2).cpu().double()
err = (got - ref).abs()
print(impl, h, w, "rel err", float(err.max() / ref.abs().max()), "finite", bool(torch.isfinite(got).all()))
if float(err.max() / ref.abs().max()) > 1e-2:
    e2 = err.amax(dim=1)[0]
    print("err by pixel (rows 0..15, cols 0..23):")
    for r in range(min(h, 16)):
        print(" ".join(f"{float(v):.1e}" for v in e2[r, :24]))
    ec = err.amax(dim=(2, 3))[0]
    print("err by channel block:", [float(ec[i * 32:(i + 1) * 32].max()) for i in range(8)])
