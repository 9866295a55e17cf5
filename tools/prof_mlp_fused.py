"""Smallest program for an ncu capture of the fused stage-0 MLP kernel (B = 64 shape)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, common
common.package()
from dgtd_b200.twig.ops.functions import texture_diffusion_func as OP
M, C = 64 * 96 * 96, 128
a = torch.randn(M, C, device="cuda").bfloat16(); w = (torch.randn(4 * C, C, device="cuda") * 0.05).bfloat16()
bias = torch.randn(4 * C, device="cuda"); col_s = torch.randn(4 * C, device="cuda"); rs = torch.rand(M, 2, device="cuda")
w2 = (torch.randn(C, 4 * C, device="cuda") * 0.05).bfloat16()
res = torch.randn(M, C, device="cuda"); g = torch.ones(C, device="cuda"); b2 = torch.zeros(C, device="cuda")
for _ in range(3):
    OP.convnext_mlp_fused_(a, rs, w, col_s, bias, w2, b2, g, res)
torch.cuda.synchronize()
print("ok")
