"""world_size-2 gloo test of the bucketed gradient all-reduce (the N>1 path of bench.py), on CPU."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from hybrid_vit_cascade_b200.dp import GradientBuckets
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.GELU(), torch.nn.Linear(32, 8), torch.nn.Linear(8, 4))
    net[2].weight.requires_grad_(False)                     # a frozen tensor must simply be skipped
    params = list(net.parameters())
    gb = GradientBuckets(params, bucket_bytes=1024)          # tiny buckets -> several all-reduces
    gb.broadcast_parameters(params)
    assert len(gb.buckets) > 1
    outs = []
    for step in range(2):
        gb.reset()
        g = torch.Generator().manual_seed(100 + 10 * step + rank)   # each rank its own shard of the batch
        x = torch.randn(5, 16, generator=g)
        net(x).square().mean().backward()
        gb.finish()
        outs.append([p.grad.clone() if p.grad is not None else None for p in params])
    # reference: average of the per-rank gradients computed locally
    ref = []
    for step in range(2):
        acc = None
        for r in range(world):
            net2 = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.GELU(), torch.nn.Linear(32, 8), torch.nn.Linear(8, 4))
            net2.load_state_dict(net.state_dict())
            g = torch.Generator().manual_seed(100 + 10 * step + r)
            net2(torch.randn(5, 16, generator=g)).square().mean().backward()
            gs = [p.grad for p in net2.parameters()]
            acc = gs if acc is None else [a + b for a, b in zip(acc, gs)]
        ref.append([a / world for a in acc])
    ok = True
    for step in range(2):
        for i, (a, b) in enumerate(zip(outs[step], ref[step])):
            if not params[i].requires_grad:
                continue
            ok = ok and torch.allclose(a, b, atol=1e-6)
    q.put((rank, ok))
    dist.destroy_process_group()


def test_bucketed_allreduce_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res
