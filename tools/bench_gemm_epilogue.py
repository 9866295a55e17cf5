"""pwconv1 at the four stage shapes: GELU vs plain bias epilogue, LN-fold vs plain (does the epilogue pace the GEMM?)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, common
common.package()
from dgtd_b200.twig.ops.functions import texture_diffusion_func as OP
def t(fn, n=20):
    for _ in range(5): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3
B = 64
for st, (hw, C) in enumerate([(96, 128), (48, 256), (24, 512), (12, 1024)]):
    M, K, N = B * hw * hw, C, 4 * C
    a = torch.randn(M, K, device="cuda").bfloat16(); w = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
    bias = torch.randn(N, device="cuda"); col_s = torch.randn(N, device="cuda"); rs = torch.rand(M, 2, device="cuda")
    h = torch.randn(M, N, device="cuda").bfloat16(); w2 = (torch.randn(K, N, device="cuda") * 0.05).bfloat16()
    res = torch.randn(M, K, device="cuda"); g = torch.ones(K, device="cuda"); b2 = torch.zeros(K, device="cuda")
    fl = 2.0 * M * N * K
    r = {}
    r["bias"] = t(lambda: OP.linear(a, w, bias, OP.ACT_NONE, OP.BF16))
    r["gelu"] = t(lambda: OP.linear(a, w, bias, OP.ACT_GELU, OP.BF16))
    r["lnfold+gelu"] = t(lambda: OP.linear_lnfold(a, w, bias, col_s, rs, OP.ACT_GELU, OP.BF16))
    r["pwconv2"] = t(lambda: OP.linear_residual_(h, w2, b2, g, None, hw * hw, res))
    if C in OP.MLP_FUSED_C:
        r["fused mlp"] = t(lambda: OP.convnext_mlp_fused_(a, rs, w, col_s, bias, w2, b2, g, res))
    print(f"stage {st} M={M} K={K} N={N}: " + "  ".join(f"{k} {v:.1f} us ({fl / v / 1e6:.0f} TF/s)" for k, v in r.items()), flush=True)
