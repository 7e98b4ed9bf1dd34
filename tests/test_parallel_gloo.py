"""CPU, world_size 2 over gloo: the multi-process host logic of the hot path -- batch sharding with
CFG pairs kept together, result gather, and the data-parallel gradient all-reduce of the trainable parameters."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, fn, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ret[rank] = fn(rank, world)
    finally:
        dist.destroy_process_group()


def _run(fn, world=2):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), fn, ret), nprocs=world, join=True)
    return [ret[r] for r in range(world)]


def _shard_job(rank, world):
    import adaface_dev_b200 as a
    n = 5                                                        # odd on purpose: ragged shards
    x = torch.arange(2 * n, dtype=torch.float32).view(2 * n, 1) * 10          # [cond_0..4, uncond_0..4]
    (local,) = a.parallel.shard_cfg_batch([x], n, rank, world)
    b, e = a.parallel.shard_range(n, rank, world)
    assert local.shape[0] == 2 * (e - b)
    cond, uncond = local.chunk(2)
    assert torch.equal(uncond - cond, torch.full_like(cond, 10.0 * n))       # pairs stayed together, order kept
    eps = a.parallel.cfg_combine(local, 4.0)                                  # stays on the rank
    full = a.parallel.gather_images(eps, n)
    return full.flatten().tolist()


def test_cfg_batch_sharding_and_gather():
    out = _run(_shard_job)
    n = 5
    x = torch.arange(2 * n, dtype=torch.float32) * 10
    ref = (x[n:] + 4.0 * (x[:n] - x[n:])).tolist()
    assert out[0] == ref and out[1] == ref


def _grad_job(rank, world):
    import adaface_dev_b200 as a
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.Linear(32, 4))     # same init on both ranks
    model[1].bias.requires_grad_(False)
    g = torch.Generator().manual_seed(100 + rank)
    x = torch.randn(8, 16, generator=g)
    model(x).square().mean().backward()
    nb = a.parallel.allreduce_gradients(model.parameters(), bucket_bytes=1024)      # tiny buckets: several collectives
    return nb, [p.grad.clone() for p in model.parameters() if p.requires_grad]


def test_gradient_allreduce_matches_single_process_mean():
    out = _run(_grad_job)
    (nb0, g0), (nb1, g1) = out
    assert nb0 == nb1 and nb0 >= 2
    for a_, b_ in zip(g0, g1):
        assert torch.equal(a_, b_)                                                   # ranks agree bit for bit
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.Linear(32, 4))
    ref = None
    for r in range(2):
        model.zero_grad()
        x = torch.randn(8, 16, generator=torch.Generator().manual_seed(100 + r))
        model(x).square().mean().backward()
        gs = [p.grad.clone() for p in list(model.parameters())[:3]]
        ref = gs if ref is None else [u + v for u, v in zip(ref, gs)]
    for got, want in zip(g0, ref):
        assert torch.allclose(got, want / 2, atol=1e-6)


def test_shard_range_covers_everything_once():
    import adaface_dev_b200 as a
    for n in (0, 1, 7, 64, 129):
        for world in (1, 2, 4, 8):
            seen = []
            for r in range(world):
                b, e = a.parallel.shard_range(n, r, world)
                seen += list(range(b, e))
            assert seen == list(range(n))
            sizes = [a.parallel.shard_range(n, r, world) for r in range(world)]
            assert max(e - b for b, e in sizes) - min(e - b for b, e in sizes) <= 1


def _none_grad_job(rank, world):
    import adaface_dev_b200 as a
    torch.manual_seed(0)
    used, one_sided, unused = (torch.nn.Parameter(torch.ones(4)) for _ in range(3))
    loss = (used * (rank + 1)).sum()
    if rank == 0:
        loss = loss + (one_sided * 3).sum()                   # gradient on rank 0 only
    loss.backward()
    a.parallel.allreduce_gradients([used, one_sided, unused])
    return (used.grad.tolist(), one_sided.grad.tolist(), unused.grad is None)


def test_gradient_allreduce_keeps_globally_unused_params_gradless():
    """A parameter without a gradient on every rank keeps .grad = None (DDP semantics: Adam must not touch it); one with a
    gradient on some rank only gets the mean with zeros from the others."""
    for used, one_sided, unused_is_none in _run(_none_grad_job):
        assert used == [1.5] * 4 and one_sided == [1.5] * 4 and unused_is_none


def _bucketer_job(rank, world):
    import adaface_dev_b200 as a
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.Linear(32, 4))     # same init on both ranks
    unused = torch.nn.Parameter(torch.ones(3))
    gb = a.parallel.GradBucketer(list(model.parameters()) + [unused], bucket_bytes=1024, expected_uses=2)
    out = []
    for it in range(2):                                                              # two iterations: zero() re-arms the buckets
        gb.zero()
        for micro in range(2):                                                       # two accumulations per iteration
            x = torch.randn(8, 16, generator=torch.Generator().manual_seed(100 * it + 10 * micro + rank))
            model(x).square().mean().backward()
        nb = gb.finish()
        out.append((nb, [p.grad.clone() for p in model.parameters()], unused.grad is None))
    gb.close()
    return out


def test_grad_bucketer_matches_mean_of_accumulated_gradients():
    r0, r1 = _run(_bucketer_job)
    for it in range(2):
        (nb0, g0, none0), (nb1, g1, none1) = r0[it], r1[it]
        assert nb0 == nb1 and nb0 >= 2 and none0 and none1
        for a_, b_ in zip(g0, g1):
            assert torch.equal(a_, b_)
        torch.manual_seed(0)
        model = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.Linear(32, 4))
        for r in range(2):
            for micro in range(2):
                x = torch.randn(8, 16, generator=torch.Generator().manual_seed(100 * it + 10 * micro + r))
                model(x).square().mean().backward()
        for got, p in zip(g0, model.parameters()):
            assert torch.allclose(got, p.grad / 2, atol=1e-6)
