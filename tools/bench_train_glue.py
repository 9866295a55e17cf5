"""Element-wise backward passes and weight-gradient GEMMs of a ConvNeXt block at the four stage shapes (B = 16)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, common
common.package()
from dgtd_b200.twig.ops.functions import train_func as TF, texture_diffusion_func as OP
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def t(fn, n=10):
    fn(); torch.cuda.synchronize(); tot = 0.0
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); tot += a.elapsed_time(b)
    return tot / n * 1e3
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
for st, (hw, C) in enumerate([(96, 128), (48, 256), (24, 512), (12, 1024)]):
    M = B * hw * hw
    g = torch.randn(M, C, device="cuda")
    dh = torch.randn(M, 4 * C, device="cuda").bfloat16(); pre = torch.randn(M, 4 * C, device="cuda").bfloat16()
    a = torch.randn(M, C, device="cuda").bfloat16()
    w1t = (torch.randn(4 * C, C, device="cuda") * 0.05).bfloat16()
    r = {}
    r["keep*g+colsum"] = t(lambda: TF.eltwise_colsum(g, 0))
    r["dH*gelu'+colsum"] = t(lambda: TF.eltwise_colsum(dh, 2, aux=pre))
    r["gelu fwd"] = t(lambda: TF.transpose_op(pre, 1, want_dst=True, want_T=False))
    r["wgrad dW1"] = t(lambda: TF.wgrad_tc_mn(dh, a))
    r["wgrad dW2"] = t(lambda: TF.wgrad_tc_mn(a, dh))
    r["dgrad da"] = t(lambda: OP.linear(dh, w1t.t().contiguous(), None, out_dtype=OP.F32))
    fl = [M * C * 6, M * 4 * C * 6, M * 4 * C * 4]
    print(f"stage {st} M={M} C={C}: " + "  ".join(f"{k} {v:.1f} us" for k, v in r.items()) +
          f"  | HBM floors {fl[0] / 6.5e6:.0f} / {fl[1] / 6.5e6:.0f} / {fl[2] / 6.5e6:.0f} us; GEMM floor {2.0 * M * C * 4 * C / 1.379e9:.0f} us", flush=True)
