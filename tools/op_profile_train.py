#!/usr/bin/env python
"""Per-entry-point timing of one fwd+bwd training step with CUDA events: op_profile_train.py [B] [fp32|bf16]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import common
TD = common.package()
from dgtd_b200.twig.ops import capi
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
prec = sys.argv[2] if len(sys.argv) > 2 else "fp32"
enc, dec = TD.build_texture_diffuser(seed=0)
enc, dec = enc.cuda().train(), dec.cuda().train()
image, depth = common.synthetic_inputs(B, 384)
image, depth = image.cuda(), depth.cuda()

def step():
    _, e3, toks = TD.texture_prompts_train(enc, dec, image, depth, precision=prec)
    loss = sum(t.mean() for row in toks for t in row) + e3.mean()
    loss.backward()
    for p in list(enc.parameters()) + list(dec.parameters()):
        p.grad = None

step()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); step(); b.record(); torch.cuda.synchronize()
print(f"wall (events) {a.elapsed_time(b):.1f} ms/step")
capi.enable_profile(True)
step()
summ = capi.profile_summary()
capi.enable_profile(False)
tot = sum(v[1] for v in summ.values())
print(f"sum of entry points: {tot:.1f} ms/step (B={B}, fwd+bwd {prec})")
for k, v in sorted(summ.items(), key=lambda kv: -kv[1][1])[:28]:
    print(f"| `{k}` | {v[0]} | {v[1]:.2f} | {100 * v[1] / tot:.1f}% |")
