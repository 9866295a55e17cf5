#!/usr/bin/env python
"""Depthwise 7x7 kernels per ConvNeXt stage: forward (TMA), with residual add, weight gradient.  bench_dwconv.py [B]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import common
TD = common.package()
from dgtd_b200.twig.ops.functions import train_func as TF

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn, n=10):
    fn(); torch.cuda.synchronize()
    tot = 0.0
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / n * 1e3


print(f"| stage | shape | fwd us | GB/s | fwd+add us | wgrad us | ln_bwd us |")
for C, h in ((128, 96), (256, 48), (512, 24), (1024, 12)):
    x = torch.randn(B, h, h, C, device="cuda")
    g = torch.randn(B, h, h, C, device="cuda")
    wT = torch.randn(49, C, device="cuda")
    b = torch.randn(C, device="cuda")
    lw = torch.randn(C, device="cuda")
    t0 = timeit(lambda: TF.dwconv7(x, wT, b))
    t1 = timeit(lambda: TF.dwconv7(x, wT, None, add=g))
    t2 = timeit(lambda: TF.dwconv7_wgrad(x, g))
    t3 = timeit(lambda: TF.ln_rows_bwd(g, x, lw, 1e-6))
    nbytes = 2 * x.numel() * 4
    print(f"| C={C} | {B}x{h}x{h} | {t0:.1f} | {nbytes / t0 / 1e3:.0f} | {t1:.1f} | {t2:.1f} | {t3:.1f} |")
