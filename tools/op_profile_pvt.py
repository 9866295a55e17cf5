#!/usr/bin/env python
"""Per-entry-point timing of pvt_v2_b2.forward_features (hot path + PVT blocks): op_profile_pvt.py [B] [bf16|fp32]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import common
TD = common.package()
from dgtd_b200.twig.ops import capi
from dgtd_b200.twig.model import pvt
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
prec = sys.argv[2] if len(sys.argv) > 2 else "bf16"
net = pvt.pvt_v2_b2().eval()
common.fill_params_(net, seed=0)
net = net.cuda()
TD.set_precision(net, prec)
image, depth = common.synthetic_inputs(B, 384)
image, depth = image.cuda(), depth.cuda()
for _ in range(2):
    net.forward_features(image, depth)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(3):
    net.forward_features(image, depth)
b.record(); torch.cuda.synchronize()
print(f"forward_features: {a.elapsed_time(b) / 3:.2f} ms/step = {B * 3 / a.elapsed_time(b) * 1e3:.0f} images/s (B={B}, {prec})")
capi.enable_profile(True)
net.forward_features(image, depth)
summ = capi.profile_summary()
capi.enable_profile(False)
tot = sum(v[1] for v in summ.values())
print(f"sum of entry points: {tot:.2f} ms/step")
for k, v in sorted(summ.items(), key=lambda kv: -kv[1][1])[:16]:
    print(f"| `{k}` | {v[0]} | {v[1]:.2f} | {100 * v[1] / tot:.1f}% |")
