#!/usr/bin/env python
"""One warm-up + one profiled pvt_v2_b2.forward_features step (for ncu --profile-from-start off)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import common
TD = common.package()
from dgtd_b200.twig.model import pvt
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
net = pvt.pvt_v2_b2().eval()
common.fill_params_(net, seed=0)
net = net.cuda()
TD.set_precision(net, "bf16")
image, depth = common.synthetic_inputs(B, 384)
image, depth = image.cuda(), depth.cuda()
net.forward_features(image, depth)
torch.cuda.synchronize()
torch.cuda.profiler.start()
net.forward_features(image, depth)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
