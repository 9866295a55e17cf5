"""Ablation timing of the fp32 W1 kernel (DGTD_ABL bit mask; results are NOT valid outputs).

The DGTD_ABL switches lived in csrc/mp_tc_f32.cu only while the experiment ran (r2, removed again: they cost 0.05 ms);
the results are recorded in profiles/r2_ncu_w1_diffusion.md.  Kept as the record of how the numbers were taken."""
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
def t(fn, n=5):
    for _ in range(2): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
for abl in [int(a) for a in sys.argv[1:]]:
    os.environ["DGTD_ABL"] = str(abl)
    print(f"abl={abl:2d}: {t(lambda: OP.message_passing_tiled(x, wgt, 1, impl='tc')):.4f} ms", flush=True)
