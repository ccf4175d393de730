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


class _TwoHeads(torch.nn.Module):
    """A trunk with two heads; forward(x, use_b=False) leaves head b trainable but unused -- ProgressiveCascadeModel(xrays,
    max_stage=1) does the same to stages 2-3 (train_progressive_4gpu.py:238 needs find_unused_parameters=True for it)."""

    def __init__(self):
        super().__init__()
        self.trunk = torch.nn.Linear(16, 16)
        self.a = torch.nn.Linear(16, 4)
        self.b = torch.nn.Linear(16, 4)

    def forward(self, x, use_b):
        h = torch.tanh(self.trunk(x))
        return self.a(h) + (self.b(h) if use_b else 0.0)


def _unused_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from hybrid_vit_cascade_b200.dp import GradientBuckets
    torch.manual_seed(0)
    net = _TwoHeads()
    params = list(net.parameters())
    gb = GradientBuckets(params, bucket_bytes=1 << 20)       # ONE bucket: used and unused parameters share it
    gb.broadcast_parameters(params)
    assert len(gb.buckets) == 1
    ok = True

    def local_grads(r, step, use_b):
        net2 = _TwoHeads()
        net2.load_state_dict(net.state_dict())
        g = torch.Generator().manual_seed(7 + 10 * step + r)
        net2(torch.randn(5, 16, generator=g), use_b).square().mean().backward()
        return [p.grad for p in net2.parameters()]

    for step, use_b in enumerate((False, True, False)):
        gb.reset()
        g = torch.Generator().manual_seed(7 + 10 * step + rank)
        net(torch.randn(5, 16, generator=g), use_b).square().mean().backward()
        gb.finish()
        per_rank = [local_grads(r, step, use_b) for r in range(world)]
        for i, p in enumerate(params):
            if per_rank[0][i] is None:                       # unused in this step: zeros in the bucket, not touched
                ok = ok and float(p.grad.abs().max()) == 0.0
                continue
            ref = sum(pr[i] for pr in per_rank) / world
            ok = ok and torch.allclose(p.grad, ref, atol=1e-6)
        n_touched = len(gb.touched(0))
        ok = ok and n_touched == (6 if use_b else 4) and gb.all_touched() == use_b
    # gradient accumulation: two micro-batches, the first under no_sync(); one all-reduce of the accumulated sum
    gb.reset()
    xs = [torch.randn(5, 16, generator=torch.Generator().manual_seed(50 + 2 * rank + k)) for k in range(2)]
    with gb.no_sync():
        net(xs[0], True).square().mean().backward()
    net(xs[1], True).square().mean().backward()
    gb.finish()
    acc = None
    for r in range(world):
        for k in range(2):
            net2 = _TwoHeads()
            net2.load_state_dict(net.state_dict())
            net2(torch.randn(5, 16, generator=torch.Generator().manual_seed(50 + 2 * r + k)), True).square().mean().backward()
            gs = [p.grad for p in net2.parameters()]
            acc = gs if acc is None else [a + b for a, b in zip(acc, gs)]
    for p, a in zip(params, acc):
        ok = ok and torch.allclose(p.grad, a / world, atol=1e-6)
    # a second synchronising backward without reset() is an error, not a silent no-op
    raised = False
    try:
        net(xs[0], True).square().mean().backward()
    except RuntimeError as e:
        raised = "reset()" in str(e)
    q.put((rank, ok and raised))
    dist.destroy_process_group()


def test_unused_trainable_parameters_and_no_sync_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_unused_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res


def _buffers_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from hybrid_vit_cascade_b200.dp import GradientBuckets
    torch.manual_seed(rank)                                  # replicas start different on purpose
    net = torch.nn.Sequential(torch.nn.Conv2d(1, 4, 3), torch.nn.BatchNorm2d(4))
    net.train()
    net(torch.randn(3, 1, 8, 8))                             # per-rank running statistics
    gb = GradientBuckets(list(net.parameters()))
    gb.broadcast_buffers(net)
    flat = torch.cat([b.double().reshape(-1) for b in net.buffers()])
    gathered = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    q.put((rank, all(torch.equal(gathered[0], t) for t in gathered) and int(net[1].num_batches_tracked) == 1))
    dist.destroy_process_group()


def test_broadcast_buffers_world2():
    """DDP(broadcast_buffers=True) semantics (train_direct_4gpu.py:146): BatchNorm running statistics follow rank 0."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_buffers_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res
