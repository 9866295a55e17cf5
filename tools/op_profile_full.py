#!/usr/bin/env python
"""Warm-cache per-entry-point timing of the full model (backbone + Hitnet decoder) with CUDA events."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import common  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--size", type=int, default=384)
ap.add_argument("--precision", default="bf16")
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--part", default="decoder", choices=["decoder", "model"])
a = ap.parse_args()
TD = common.package()
from dgtd_b200.twig.model import hitnet  # noqa: E402
from dgtd_b200.twig.ops import capi  # noqa: E402
net = hitnet.Hitnet().eval()
common.hitnet_fixture_params_(net, seed=0)
net = net.cuda()
TD.set_precision(net, a.precision)
image, depth = common.synthetic_inputs(a.batch, a.size)
image, depth = image.cuda(), depth.cuda()
_, feats = net.backbone._forward_features_nhwc(image, depth)
fn = (lambda: net.decode(feats, want_stage_preds=False)) if a.part == "decoder" else (lambda: net.predict_logits(image, depth))
for _ in range(2):
    fn()
torch.cuda.synchronize()
capi.enable_profile(True)
for _ in range(a.steps):
    fn()
summ = capi.profile_summary()
capi.enable_profile(False)
tot = sum(v[1] for v in summ.values()) / a.steps
print(f"{a.part}: sum of entry points {tot:.3f} ms/step (B={a.batch}, {a.size}x{a.size}, {a.precision})")
print("| entry point | calls/step | ms/step | share |\n|---|---|---|---|")
for k, v in sorted(summ.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k}` | {v[0] // a.steps} | {v[1] / a.steps:.3f} | {100 * v[1] / a.steps / tot:.1f}% |")
