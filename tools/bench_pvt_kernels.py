#!/usr/bin/env python
"""GPU time of the non-GEMM PVT-v2 kernels per stage shape (CUDA events, L2 flushed between launches):
bench_pvt_kernels.py [B] [fwd|bwd|attn|all]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import common
common.package()
from dgtd_b200.twig.ops.functions import pvt_func as PF, pvt_train_func as PT
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
what = sys.argv[2] if len(sys.argv) > 2 else "all"
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
STAGES = [(96, 512, 1, 3), (48, 1024, 2, 4), (24, 1280, 5, 6), (12, 2048, 8, 3)]   # (grid, hidden, heads, blocks)

def timed(fn, n=5):
    fn(); torch.cuda.synchronize()
    tot = 0.0
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / n

g = torch.Generator(device="cuda").manual_seed(0)
total = {}
for s, (hw, C, heads, nblk) in enumerate(STAGES):
    x = torch.randn(B, hw, hw, C, device="cuda", generator=g).to(torch.bfloat16)
    wT = torch.randn(9, C, device="cuda", generator=g) * 0.3
    bias = torch.randn(C, device="cuda", generator=g) * 0.1
    go = torch.randn(B, hw, hw, C, device="cuda", generator=g)
    n = x.numel()
    if what in ("fwd", "all"):
        ms = timed(lambda: PF.dwconv3_gelu(x, wT, bias))
        print(f"stage {s + 1} dwconv3_gelu fwd bf16 {hw}x{hw}x{C} B={B}: {ms * 1e3:.1f} us, {n * 4 / ms / 1e6:.0f} GB/s (4 B/elt)")
        total["fwd"] = total.get("fwd", 0) + ms * nblk
    if what in ("bwd", "all"):
        ms = timed(lambda: PT.dwconv3_gelu_bwd(x, wT, bias, go))
        print(f"stage {s + 1} dwconv3_gelu bwd (du + tap grads + dgrad) : {ms * 1e3:.1f} us, {n * 18 / ms / 1e6:.0f} GB/s (18 B/elt)")
        total["bwd"] = total.get("bwd", 0) + ms * nblk
    if what in ("attn", "all"):
        N, Nk = hw * hw, 144
        Ca = heads * 64
        q = torch.randn(B * N, Ca, device="cuda", generator=g).to(torch.bfloat16)
        kv = torch.randn(B * Nk, 2 * Ca, device="cuda", generator=g).to(torch.bfloat16)
        do = torch.randn(B * N, Ca, device="cuda", generator=g)
        o = PF.attention(q, kv, B, N, Nk, heads)
        ms = timed(lambda: PF.attention(q, kv, B, N, Nk, heads))
        fl = 4.0 * B * heads * N * Nk * 64
        print(f"stage {s + 1} attention fwd N={N} Nk={Nk} heads={heads}: {ms * 1e3:.1f} us, {fl / ms / 1e9:.1f} TFLOP/s")
        total["attn_fwd"] = total.get("attn_fwd", 0) + ms * nblk
        ms = timed(lambda: PT.attention_bwd(q, kv, o, do, B, N, Nk, heads))
        print(f"stage {s + 1} attention bwd: {ms * 1e3:.1f} us, {2.5 * fl / ms / 1e9:.1f} TFLOP/s")
        total["attn_bwd"] = total.get("attn_bwd", 0) + ms * nblk
    del x, go
print({k: round(v, 3) for k, v in total.items()}, "ms per model pass (x blocks per stage)")
