#!/usr/bin/env python
"""Per-entry-point timing of one TRAINING step of the whole model (`cod.forward(mode='loss')` + backward; backbone
with the texture prompts, Hitnet decoder with train-mode BatchNorm, deep-supervision loss):
op_profile_full_train.py [B] [fp32|bf16] [S]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import common
common.package()
from dgtd_b200.twig.ops import capi
from dgtd_b200.twig.model import hitnet
from dgtd_b200.twig.model.texture_diffuser import set_precision
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
prec = sys.argv[2] if len(sys.argv) > 2 else "bf16"
S = int(sys.argv[3]) if len(sys.argv) > 3 else 384
net = hitnet.cod(win_size=22, filter_ratio=0.9, using_sam=True, using_depth=True, finetune=True, binary_thresh=0.2)
common.hitnet_fixture_params_(net.hitnet, seed=0)
net = net.cuda().train()
set_precision(net, prec)
image, depth = common.synthetic_inputs(B, S)
_, label = common.loss_inputs(B, S, S, seed=11)
image, depth, label = image.cuda(), depth.cuda(), label.cuda()

def step():
    loss = net(None, image, label, depth, mode="loss")["loss"]
    loss.backward()
    for p in net.parameters():
        p.grad = None
    return loss

for _ in range(2):
    step()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 3
a.record()
for _ in range(n):
    step()
b.record(); torch.cuda.synchronize()
print(f"wall (events) {a.elapsed_time(b) / n:.1f} ms/step; peak memory {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
l0 = capi.launch_count()
step()
print("launches per step:", capi.launch_count() - l0)
capi.enable_profile(True)
step()
summ = capi.profile_summary()
capi.enable_profile(False)
tot = sum(v[1] for v in summ.values())
print(f"sum of entry points: {tot:.1f} ms/step (B={B}, {S}x{S}, fwd+bwd {prec})")
for k, v in sorted(summ.items(), key=lambda kv: -kv[1][1])[:36]:
    print(f"| `{k}` | {v[0]} | {v[1]:.2f} | {100 * v[1] / tot:.1f}% |")
