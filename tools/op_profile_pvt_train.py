#!/usr/bin/env python
"""Per-entry-point timing of one fwd+bwd step of the PVT-v2 backbone + hot path (forward_features under autograd):
op_profile_pvt_train.py [B] [fp32|bf16] [S]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import common
common.package()
from dgtd_b200.twig.ops import capi
from dgtd_b200.twig.model import pvt
from dgtd_b200.twig.model.texture_diffuser import set_precision
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
prec = sys.argv[2] if len(sys.argv) > 2 else "bf16"
S = int(sys.argv[3]) if len(sys.argv) > 3 else 384
net = pvt.pvt_v2_b2()
common.fill_params_(net, seed=0)
net = net.cuda().train()
set_precision(net, prec)
image, depth = common.synthetic_inputs(B, S)
image, depth = image.cuda(), depth.cuda()

def step():
    _, outs = net.forward_features(image, depth)
    loss = sum(o.mean() for o in outs)
    loss.backward()
    for p in net.parameters():
        p.grad = None

step()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); step(); b.record(); torch.cuda.synchronize()
print(f"wall (events) {a.elapsed_time(b):.1f} ms/step; peak memory {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
capi.enable_profile(True)
step()
summ = capi.profile_summary()
capi.enable_profile(False)
tot = sum(v[1] for v in summ.values())
print(f"sum of entry points: {tot:.1f} ms/step (B={B}, {S}x{S}, fwd+bwd {prec})")
for k, v in sorted(summ.items(), key=lambda kv: -kv[1][1])[:32]:
    print(f"| `{k}` | {v[0]} | {v[1]:.2f} | {100 * v[1] / tot:.1f}% |")
