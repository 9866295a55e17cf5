import os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import torch, common
common.package()
from dgtd_b200.twig.ops.functions import pvt_func as PF
B, hw, C = 64, 96, 512
x = torch.randn(B, hw, hw, C, device="cuda").to(torch.bfloat16)
wT = torch.randn(9, C, device="cuda") * 0.3
bias = torch.randn(C, device="cuda") * 0.1
for _ in range(2):
    PF.dwconv3_gelu(x, wT, bias)
torch.cuda.synchronize()
