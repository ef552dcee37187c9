"""CPU, world_size 2, gloo: the N > 1 host logic (batch sharding, flattened gradient all-reduce, loss reduction)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from rdfc_gan_b200.parallel import ShardedGenerator, allreduce_gradients, reduce_losses, shard_bounds

        class Fake:     # stands in for the CUDA generator: tags every image with its global index
            def __call__(self, rgb, depth, normal):
                return {"pred_depth": depth * 2, "confidence_map_2": normal[:, :1] + 1}
        B = 5
        depth = torch.arange(B, dtype=torch.float32).view(B, 1, 1, 1).expand(B, 1, 2, 3).contiguous()
        rgb, normal = torch.zeros(B, 3, 2, 3), torch.ones(B, 3, 2, 3)
        lo, hi = shard_bounds(B, world, rank)
        local = ShardedGenerator(Fake())(rgb, depth, normal)
        assert local["pred_depth"].shape[0] == hi - lo
        full = ShardedGenerator(Fake())(rgb, depth, normal, gather=True)
        assert torch.equal(full["pred_depth"], depth * 2) and full["confidence_map_2"].shape == (B, 1, 2, 3)

        torch.manual_seed(0)
        net = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.Linear(3, 2), torch.nn.Linear(2, 2))
        for p in net[2].parameters():
            p.requires_grad_(False)          # like fuse_layer5 / the frozen NLSPN dummies: never gets a grad
        x = torch.full((2, 4), float(rank + 1))
        net[1](net[0](x)).sum().backward()
        mine = [p.grad.clone() for p in net.parameters() if p.grad is not None]
        n = allreduce_gradients(net.parameters())
        assert n == sum(g.numel() for g in mine) == 4 * 3 + 3 + 3 * 2 + 2
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        for i, p in enumerate(p for p in net.parameters() if p.grad is not None):
            expect = sum(g[i] for g in gathered) / world
            assert torch.allclose(p.grad, expect, atol=1e-6)
        # GradientBucket: grads are views of one flat buffer; a parameter that got no grad on ONE rank must not desynchronise
        from rdfc_gan_b200.parallel import GradientBucket
        torch.manual_seed(1)
        net2 = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.Linear(3, 2))
        bucket = GradientBucket(net2.parameters())
        for step in range(2):
            bucket.zero()
            x = torch.full((2, 4), float(rank + 1 + step))
            y = net2[0](x) if rank == 0 else net2[1](net2[0](x))          # rank 0 never touches net2[1] ("unused branch")
            y.sum().backward()
            local = [p.grad.clone() for p in net2.parameters()]
            assert all(p.grad.data_ptr() >= bucket.flat[(p.dtype, p.device)].data_ptr() for p in net2.parameters())
            assert bucket.allreduce() == 4 * 3 + 3 + 3 * 2 + 2
            gathered = [None] * world
            dist.all_gather_object(gathered, local)
            for i, p in enumerate(net2.parameters()):
                assert torch.allclose(p.grad, sum(g[i] for g in gathered) / world, atol=1e-6), (step, i)
        if rank == 1:                                   # world 2, one image: rank 1 owns nothing and must say so
            with pytest.raises(RuntimeError):
                ShardedGenerator(Fake())(rgb[:1], depth[:1], normal[:1])
        red = reduce_losses({"loss_G": torch.tensor(float(rank)), "loss_D": 2.0})
        assert abs(red["loss_G"] - 0.5) < 1e-6 and abs(red["loss_D"] - 2.0) < 1e-6
        q.put((rank, "ok"))
    except Exception as e:      # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_shard_bounds_cover_batch():
    from rdfc_gan_b200.parallel import shard_bounds
    for B in (1, 5, 32, 256, 257):
        for world in (1, 2, 4, 8):
            spans = [shard_bounds(B, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
    with pytest.raises(ValueError):
        shard_bounds(8, 2, 2)


def test_two_rank_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
