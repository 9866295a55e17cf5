"""GPU time of the B = 16 training GEMM shapes, measured as a captured graph of 20 back-to-back launches (no host gaps)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, common
common.package()
from dgtd_b200.twig.ops.functions import train_func as TF, texture_diffusion_func as OP
def graph_time(fn, reps=20):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3): fn()
        s.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(reps): fn()
        g.replay(); s.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(s)
        for _ in range(5): g.replay()
        b.record(s); s.synchronize()
    return a.elapsed_time(b) / (5 * reps) * 1e3
B = 16
for st, (hw, C) in enumerate([(96, 128), (48, 256), (24, 512), (12, 1024)]):
    M = B * hw * hw
    a = torch.randn(M, C, device="cuda").bfloat16(); h = torch.randn(M, 4 * C, device="cuda").bfloat16()
    w1 = (torch.randn(4 * C, C, device="cuda") * 0.05).bfloat16(); w1t = w1.t().contiguous()
    w2 = (torch.randn(C, 4 * C, device="cuda") * 0.05).bfloat16(); w2t = w2.t().contiguous()
    b1 = torch.randn(4 * C, device="cuda")
    o1 = torch.empty(M, 4 * C, device="cuda", dtype=torch.bfloat16); o2 = torch.empty(M, C, device="cuda")
    r = {}
    r["fwd pwconv1 (bf16 out)"] = graph_time(lambda: OP.linear(a, w1, b1, out=o1))
    r["dgrad dh = g @ W2 (bf16 out)"] = graph_time(lambda: OP.linear(a, w2t, None, out=o1))
    r["dgrad da = dh @ W1 (fp32 out)"] = graph_time(lambda: OP.linear(h, w1t, None, out_dtype=OP.F32, out=o2))
    r["wgrad dW1"] = graph_time(lambda: TF.wgrad_tc_mn(h, a))
    r["wgrad dW2"] = graph_time(lambda: TF.wgrad_tc_mn(a, h))
    print(f"stage {st} M={M} C={C} (floor {2.0 * M * C * 4 * C / 1.379e9:.0f} us): " + "  ".join(f"{k} {v:.1f}" for k, v in r.items()), flush=True)
