#!/usr/bin/env python
"""Stage-2 depthwise 7x7 (+ LN rows) kernels alone, for an `ncu --set full` capture: profile_dwconv.py [B]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import common
TD = common.package()
from dgtd_b200.twig.ops.functions import texture_diffusion_func as OP
from dgtd_b200.twig.ops.capi import BF16

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
C, h = 512, 24
x = torch.randn(B, h, h, C, device="cuda")
wT = torch.randn(49, C, device="cuda")
b = torch.randn(C, device="cuda")
lw, lb = torch.randn(C, device="cuda"), torch.randn(C, device="cuda")
ws = torch.empty_like(x)
OP.dwconv7_ln_tma(x, wT, b, lw, lb, BF16, ws)
torch.cuda.synchronize()
torch.cuda.profiler.start()
OP.dwconv7_ln_tma(x, wT, b, lw, lb, BF16, ws)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
