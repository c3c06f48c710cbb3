"""Data-parallel host logic on CPU: world_size-2 gloo run of GradBucketReducer (static live set, persistent flat
buckets fired from grad hooks, gradients re-pointed at the reduced buffers, a member without a gradient, summed vs
averaged results) and rank-sharding of the sprite loader."""
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
    from lunaris_orion_b200.data import SpriteLoader, SyntheticSprites
    from lunaris_orion_b200.train_hybrid import GradBucketReducer
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(64, 300), torch.nn.ReLU(), torch.nn.Linear(300, 200),
                              torch.nn.Linear(200, 8))
    dead = torch.nn.Parameter(torch.zeros(5))            # in the live set but without a gradient on this step
    params = list(net.parameters()) + [dead]
    red = GradBucketReducer(params, bucket_mb=0.1)
    assert len(red.buckets) > 1
    red.keep_local = True
    x = torch.full((4, 64), float(rank + 1))
    net(x).pow(2).sum().backward()
    fired_from_hooks = len(red.inflight)
    red.finish(average=True)
    ok = fired_from_hooks >= len(red.buckets) - 1        # only the bucket holding `dead` waits for finish()
    # p.grad now IS a slice of the bucket's flat buffer holding the average of the per-rank local gradients
    for i, b in enumerate(red.buckets):
        gathered = [torch.empty_like(red.local[i]) for _ in range(world)]
        dist.all_gather(gathered, red.local[i])
        mean = sum(gathered) / world
        ok = ok and torch.allclose(red.flat[i], mean, rtol=1e-5, atol=1e-6)
        for p in b:
            if p is dead:
                continue
            ok = ok and p.grad.data_ptr() == red.views[id(p)].data_ptr() and p.grad.shape == p.shape
    ok = ok and dead.grad is None
    # second step: hooks re-arm, sums (not averages) when the optimizer folds 1/world in
    net.zero_grad(set_to_none=True)
    net(x * 2).sum().backward()
    local = [p.grad.clone() for p in net.parameters()]
    red.finish()
    tot = [g.clone() for g in local]
    for t in tot:
        dist.all_reduce(t)
    ok = ok and all(torch.allclose(p.grad, t, rtol=1e-5, atol=1e-6) for p, t in zip(net.parameters(), tot))
    # disabled (inside an accumulation window): gradients stay local
    red.enabled = False
    net.zero_grad(set_to_none=True)
    net(x).sum().backward()
    ok = ok and not red.inflight and all(p.grad.data_ptr() != red.views[id(p)].data_ptr() for p in net.parameters())
    ds = SyntheticSprites(16)
    ld = SpriteLoader(ds, list(range(16)), 4, "cpu", rank=rank, world=world, generator=torch.Generator().manual_seed(1))
    first, idx = next(ld.epoch(with_indices=True))
    ok = ok and first.shape == (4, 3, 128, 128) and float(first.min()) >= -1.0 and float(first.max()) <= 1.0
    all_idx = [None] * world
    dist.all_gather_object(all_idx, idx.tolist())
    ok = ok and len(set(all_idx[0]) & set(all_idx[1])) == 0
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
