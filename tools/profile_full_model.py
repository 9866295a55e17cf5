#!/usr/bin/env python
"""One warm-up + one profiled `cod.forward(mode='predict')` step + metrics (for ncu --profile-from-start off)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import common  # noqa: E402

TD = common.package()
from dgtd_b200.twig.metric import sod_metrics  # noqa: E402
from dgtd_b200.twig.model import hitnet  # noqa: E402
from dgtd_b200.twig.ops import capi  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
net = hitnet.cod(binary_thresh=0.2).eval()
common.hitnet_fixture_params_(net.hitnet, seed=0)
net = net.cuda()
TD.set_precision(net, "bf16")
image, depth = common.synthetic_inputs(B, 384)
image, depth = image.cuda(), depth.cuda()
label = (torch.rand(B, 1, 384, 384) > 0.5).float().cuda()
prob, _ = net(None, image, label, depth, mode="predict")
sod_metrics(prob, label)
torch.cuda.synchronize()
n0 = capi.launch_count()
torch.cuda.profiler.start()
prob, _ = net(None, image, label, depth, mode="predict")
sod_metrics(prob, label)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("launches per step:", capi.launch_count() - n0)
