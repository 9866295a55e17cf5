"""Probe: does an NCCL all-reduce forked inside a captured CUDA graph overlap the compute branch on this box?"""
import os, sys, torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
os.environ.setdefault("NCCL_DEBUG", "WARN")
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
a = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
b = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
big = torch.randn(2048, 4096, 128, device="cuda")      # 4 GB elementwise = memory-bound branch
g = torch.zeros(88_000_000, device="cuda")

def compute(kind):
    if kind == "mm":
        c = a
        for _ in range(8):
            c = c @ b
        return c
    x = big
    for _ in range(4):
        x = x * 1.0001
    return x

def timed(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); dist.barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n

def capture(kind, mode, nb=12):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    def body():
        works = []
        if mode == "fork":
            n = g.numel() // nb
            for i in range(nb):
                works.append(dist.all_reduce(g[i * n:(i + 1) * n], op=dist.ReduceOp.AVG, async_op=True))
        out = compute(kind)
        if mode == "after":
            works.append(dist.all_reduce(g, op=dist.ReduceOp.AVG, async_op=True))
        for w in works: w.wait()
        return out
    with torch.cuda.stream(side):
        for _ in range(3): body()
    torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr, capture_error_mode="thread_local"):
        body()
    return gr

for kind in ("mm", "ew"):
    res = {}
    for mode in ("none", "after", "fork"):
        gr = capture(kind, mode)
        res[mode] = timed(gr.replay)
        del gr
    ar = timed(lambda: dist.all_reduce(g, op=dist.ReduceOp.AVG))
    if rank == 0:
        print(f"{kind}: compute only {res['none']:.3f} ms, + all-reduce after {res['after']:.3f}, forked at start (12 buckets) {res['fork']:.3f}; standalone all-reduce {ar:.3f} ms")
dist.destroy_process_group()
