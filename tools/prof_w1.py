"""Smallest program for an ncu capture of the W1 tensor-pipe diffusion kernel (configs[3] size)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, common
common.package()
from dgtd_b200.twig.ops.functions import texture_diffusion_func as OP
S, C = 1024, 256
g = torch.Generator().manual_seed(0)
dt = torch.bfloat16 if (len(sys.argv) < 2 or sys.argv[1] == "bf16") else torch.float32
x = torch.randn(1, S, S, C, generator=g).to(dt).cuda()
wgt = torch.rand(1, 49, S, S, generator=g).cuda()
for _ in range(4):
    OP.message_passing_tiled(x, wgt, 1)
torch.cuda.synchronize()
print("ok")
