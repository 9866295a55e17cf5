"""N>1 host logic on CPU: world_size-2 gloo processes shard a batch by image with no data-path
collective and agree on the max-over-ranks time (the bench contract)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import common


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    common.package()
    from dgtd_b200.twig import sharding
    lo, hi = sharding.shard_range(total, world, rank)
    # every rank regenerates only its own images (seed = global image index)
    imgs = [common.synthetic_inputs(1, 16, seed=i)[0] for i in range(lo, hi)]
    checksum = float(sum(float(t.double().sum()) for t in imgs))
    t_max = sharding.max_over_ranks(0.5 + rank, torch.device("cpu"))
    n_sum = sharding.sum_over_ranks(hi - lo, torch.device("cpu"))
    q.put((rank, lo, hi, checksum, t_max, n_sum))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_by_image():
    world, total = 2, 7
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, lo0, hi0, c0, t0, n0), (r1, lo1, hi1, c1, t1, n1) = res
    assert (lo0, hi0, lo1, hi1) == (0, 4, 4, 7)                    # disjoint cover, sizes differ by <= 1
    assert t0 == t1 == 1.5 and n0 == n1 == total                  # max / sum over ranks agree everywhere
    ref = sum(float(common.synthetic_inputs(1, 16, seed=i)[0].double().sum()) for i in range(total))
    assert abs((c0 + c1) - ref) < 1e-9                             # the union of shards is the whole batch


def test_shard_ranges_cover_without_overlap():
    common.package()
    from dgtd_b200.twig import sharding
    for total in (0, 1, 8, 13, 64):
        for world in (1, 2, 4, 8):
            spans = [sharding.shard_range(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = sharding.shard_sizes(total, world)
            assert max(sizes) - min(sizes) <= 1 and sum(sizes) == total
