"""ncu target: two launches each of the bf16- and fp32-storage W1 diffusion kernels at configs[3] size."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, common
common.package()
from dgtd_b200.twig.ops.functions import texture_diffusion_func as OP
S, C = 1024, 256
g = torch.Generator().manual_seed(0)
x = torch.randn(1, S, S, C, generator=g).cuda()
wgt = torch.rand(1, 49, S, S, generator=g).cuda()
xb = x.to(torch.bfloat16)
for _ in range(3):
    OP.message_passing_tiled(xb, wgt, 1)
    OP.message_passing_tiled(x, wgt, 1)
torch.cuda.synchronize()
print("ok")
