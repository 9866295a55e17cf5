#!/usr/bin/env python
"""One warm-up + N steps of the hot path, for ncu (launch list / --set full captures)."""
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
ap.add_argument("--steps", type=int, default=1)
ap.add_argument("--precision", default="bf16")
a = ap.parse_args()
TD = common.package()
from dgtd_b200.twig.ops import capi  # noqa: E402
enc, dec = TD.build_texture_diffuser(seed=0)
enc, dec = enc.cuda().eval(), dec.cuda().eval()
image, depth = common.synthetic_inputs(a.batch, a.size)
image, depth = image.cuda(), depth.cuda()
TD.texture_prompts(enc, dec, image, depth, precision=a.precision, want_embedding3=False)
torch.cuda.synchronize()
n0 = capi.launch_count()
torch.cuda.profiler.start()      # ncu --profile-from-start off: capture only the steady-state step(s)
for _ in range(a.steps):
    TD.texture_prompts(enc, dec, image, depth, precision=a.precision, want_embedding3=False)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("launches per step:", (capi.launch_count() - n0) // a.steps)
