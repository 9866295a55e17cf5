#!/usr/bin/env python
"""Two steps of the tiled MessagePassing kernel at microbench scale (for ncu)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import common
common.package()
from dgtd_b200.twig.ops.functions import texture_diffusion_func as OP
S, C = 1024, 256
x = torch.randn(1, S, S, C, device="cuda")
w = torch.rand(1, 49, S, S, device="cuda")
OP.message_passing_tiled(x, w, 1)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); OP.message_passing_tiled(x, w, 2); b.record(); torch.cuda.synchronize()
print("ms per step", a.elapsed_time(b) / 2)
