"""Multi-GPU check of the bucketed in-graph gradient reduction (run under torchrun, one rank per GPU):
the gradients after a replay with reduce="bucketed" must equal (a) the round-1 scheme reduce="after" and (b) the
mean over ranks of the local gradients (reduce="none" + an explicit all-reduce), bit for bit up to the summation
order inside NCCL (tolerance 1e-6 relative)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, torch.distributed as dist
import common

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
os.environ.setdefault("NCCL_DEBUG", "WARN")
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
TD = common.package()
from dgtd_b200.twig import graphs
S, B = 192, 2
enc, dec = TD.build_texture_diffuser(seed=0)
common.perturb_regressor_(enc)
enc, dec = enc.cuda().train(), dec.cuda().train()
for m in list(enc.modules()) + list(dec.modules()):
    if isinstance(m, TD.DropPath):
        m.drop_prob = 0.0                      # deterministic step
image, depth = common.synthetic_inputs(B, S, seed=50 + rank)     # different images per rank
image, depth = image.cuda(), depth.cuda()
out = {}
for mode in ("none", "after", "bucketed"):
    step = graphs.GraphedTrainStep(enc, dec, image, depth, precision="bf16", reduce=mode, bucket_bytes=25 << 20)
    step()
    step()
    torch.cuda.synchronize()
    g = step.flat_grad.clone()
    if mode == "none":
        dist.all_reduce(g, op=dist.ReduceOp.AVG)
    out[mode] = g
    if mode == "bucketed":
        desc, order = step.bucketer.describe(), list(step.bucketer.launch_order)
    step.close()
    del step
ref = out["none"]
scale = float(ref.abs().max())
e1 = float((out["after"] - ref).abs().max()) / scale
e2 = float((out["bucketed"] - ref).abs().max()) / scale
if rank == 0:
    print(f"world {world}: |after - mean(local)| = {e1:.2e}, |bucketed - mean(local)| = {e2:.2e} (relative to max|grad| {scale:.3e})")
    print(desc, "; launch order", order)
assert e1 <= 1e-6 and e2 <= 1e-6, (e1, e2)
dist.barrier()
dist.destroy_process_group()
if rank == 0:
    print("OK")
