"""world_size-2 gloo tests (CPU) of the data-parallel plumbing: the flat gradient bucket (C1) and the smartaverage
sum exchange (C2).  The arithmetic kernels are GPU-only; what is tested here is the host-side sharding/reduction
logic that bench.py --gpus N and style_grams_smartaverage(group=...) rely on."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, fn, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        out[rank] = fn(rank, world)
    finally:
        dist.destroy_process_group()


def _run(fn, world=2):
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), fn, out), nprocs=world, join=True)
    return dict(out)


def _grad_job(rank, world):
    from artist_style_transfer_b200 import StyleTransfer, dp
    torch.manual_seed(0)
    net = StyleTransfer(device="cpu")
    params = list(net.parameters())
    g = torch.Generator().manual_seed(100 + rank)
    for p in params:
        p.grad = torch.randn(p.shape, generator=g)
    expect = []
    for p in params:           # what the average over both ranks must be
        acc = torch.zeros_like(p)
        for r in range(world):
            pass
        expect.append(acc)
    bucket = dp.GradBucket()
    bucket.allreduce_mean(params, None)
    assert bucket.flat.numel() == 1712771                    # one flat bucket, SURVEY C1
    return [p.grad.clone() for p in params[:4]] + [params[-1].grad.clone()]


def test_grad_bucket_allreduce_mean():
    res = _run(_grad_job)
    # both ranks hold identical averaged gradients equal to the mean of the two seeded streams
    for a, b in zip(res[0], res[1]):
        assert torch.equal(a, b)
    from artist_style_transfer_b200 import StyleTransfer
    torch.manual_seed(0)
    params = list(StyleTransfer(device="cpu").parameters())
    gens = [torch.Generator().manual_seed(100 + r) for r in range(2)]
    per_rank = [[torch.randn(p.shape, generator=g) for p in params] for g in gens]
    want = [(per_rank[0][i] + per_rank[1][i]) / 2 for i in range(len(params))]
    idx = [0, 1, 2, 3, len(params) - 1]
    for got, i in zip(res[0], idx):
        torch.testing.assert_close(got, want[i], rtol=1e-6, atol=1e-7)


def _arena_job(rank, world):
    """The fused-optimizer path's exchange: replicas that start DIFFERENT are made identical by the trainer's broadcast,
    and the flat gradient arena (tap-major segments, read back through the strided p.grad views) is averaged in place."""
    from artist_style_transfer_b200 import StyleTransfer, arena as arena_mod, dp
    torch.manual_seed(rank)                                   # deliberately different replicas
    net = StyleTransfer(device="cpu", precision="fast")
    params = list(net.parameters())
    before = params[4].detach().clone()
    dp.broadcast_parameters(params, None)
    ar = arena_mod.TransferArena(net._stages(), "fast", torch.device("cpu"))
    gbuf = ar.new_grad_buffer()
    g = torch.Generator().manual_seed(200 + rank)
    vals = []
    for view, p in zip(ar.grad_views(gbuf), ar.params()):
        r = torch.randn(p.shape, generator=g)
        view.copy_(r)                                         # strided write into the arena
        vals.append(r)
    dp.allreduce_mean_flat(gbuf, None)
    views = ar.grad_views(gbuf)
    return {"p3_before": before, "p3_after": params[4].detach().clone(), "vals": [vals[0], vals[5], vals[-2]],
            "avg": [views[0].clone(), views[5].clone(), views[-2].clone()], "numel": ar.g_numel}


def test_arena_allreduce_and_parameter_broadcast():
    res = _run(_arena_job)
    assert not torch.equal(res[0]["p3_before"], res[1]["p3_before"])           # the replicas really differed
    assert torch.equal(res[0]["p3_after"], res[1]["p3_after"]) and torch.equal(res[0]["p3_after"], res[0]["p3_before"])
    for i in range(3):
        want = (res[0]["vals"][i] + res[1]["vals"][i]) / 2
        torch.testing.assert_close(res[0]["avg"][i], want, rtol=1e-6, atol=1e-7)
        assert torch.equal(res[0]["avg"][i], res[1]["avg"][i])
    assert res[0]["numel"] >= 1712771                                           # every parameter has a segment


def _sum_job(rank, world):
    from artist_style_transfer_b200 import dp
    n = 11
    lo, hi = dp.shard_range(n, rank, world)
    feats = [torch.full((2, 3), float(i)) for i in range(n)]          # "feature maps" of 11 paintings
    acc = torch.zeros(2, 3)
    for i in range(lo, hi):
        acc += feats[i]
    count = torch.tensor([float(hi - lo)])
    dp.allreduce_sums([acc, count], None)
    return (lo, hi, acc.clone(), float(count))


def test_smartaverage_sharded_sum():
    res = _run(_sum_job)
    ranges = sorted((res[r][0], res[r][1]) for r in res)
    assert ranges == [(0, 6), (6, 11)]                                 # contiguous, sizes differ by <= 1
    for r in res:
        assert res[r][3] == 11.0
        torch.testing.assert_close(res[r][2], torch.full((2, 3), float(sum(range(11)))))


def test_shard_range_covers_everything():
    from artist_style_transfer_b200 import dp
    for n in (0, 1, 7, 8, 4096):
        for w in (1, 2, 3, 8):
            parts = [dp.shard_range(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
            assert max(b - a for a, b in parts) - min(b - a for a, b in parts) <= 1
