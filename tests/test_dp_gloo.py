"""Data-parallel host logic on CPU: world_size-2 gloo run of GradBucketReducer (bucketing, None-set members, averaging)
and rank-sharding of SpriteData."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lunaris_orion_b200.train_hybrid import GradBucketReducer, SpriteData
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(64, 300), torch.nn.ReLU(), torch.nn.Linear(300, 200),
                              torch.nn.Linear(200, 8))
    dead = torch.nn.Parameter(torch.zeros(5))            # never receives a gradient (reference None-set)
    params = list(net.parameters()) + [dead]
    red = GradBucketReducer(params, bucket_mb=0.1)
    assert len(red.buckets) > 1
    x = torch.full((4, 64), float(rank + 1))
    net(x).pow(2).sum().backward()
    local = [p.grad.clone() for p in net.parameters()]
    red.finish()
    gathered = [None] * world
    dist.all_gather_object(gathered, [g.tolist() for g in local])
    ok = dead.grad is None
    for i, p in enumerate(net.parameters()):
        mean = sum(torch.tensor(g[i]) for g in gathered) / world
        ok = ok and torch.allclose(p.grad, mean, rtol=1e-5, atol=1e-6)
    # second step: hooks re-arm
    net.zero_grad(set_to_none=True)
    net(x * 2).sum().backward()
    red.finish()
    ok = ok and all(p.grad is not None for p in net.parameters())
    d = SpriteData("synthetic", 4, rank, world)
    first = next(d.batches(0, torch.device("cpu")))
    ok = ok and first.shape == (4, 3, 128, 128) and float(first.min()) >= -1.0 and float(first.max()) <= 1.0
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_bucket_reducer_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
