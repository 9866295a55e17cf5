"""Per-role timeline of one CTA of the fp32 W1 kernel (debug build with the TR() probes).

The TR() probes (clock64 + event id into a global buffer, block 3 only, DGTD_TRACE_PTR) lived in csrc/mp_tc_f32.cu only
while the experiment ran (r2); results in profiles/r2_ncu_w1_diffusion.md.  Kept as the record of how the numbers were taken."""
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
for _ in range(3): OP.message_passing_tiled(x, wgt, 1, impl="tc")
torch.cuda.synchronize()
tr = torch.zeros(9 * 4096, dtype=torch.int64, device="cuda")
os.environ["DGTD_TRACE_PTR"] = str(tr.data_ptr())
OP.message_passing_tiled(x, wgt, 1, impl="tc")
torch.cuda.synchronize()
import numpy as np
np.save(os.path.join(ROOT, "gpurun_out", "trace_w1.npy"), tr.cpu().numpy().reshape(9, 4096))
print("saved")
