"""Multi-GPU host logic: one process per GPU (torch.distributed), batch-sharded inference with NO data-path
collective, and the single flattened gradient all-reduce that replaces DistributedDataParallel for the training step
(C/lib/models/rdfc_gan.py:102-119 wraps the nets in DDP; C/lib/models/base.py:121-132 all-reduces one loss scalar per
key).  Works with any backend: NCCL over NVLink on the B200 box, gloo in the CPU tests."""
import torch
import torch.distributed as dist


def shard_bounds(batch, world, rank):
    """Contiguous slice [lo, hi) of a global batch owned by `rank` (remainder images go to the first ranks)."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world {world}")
    base, rem = divmod(batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _world():
    return (dist.get_world_size(), dist.get_rank()) if dist.is_available() and dist.is_initialized() else (1, 0)


class ShardedGenerator:
    """Runs `generator` on this rank's slice of a global batch.  Inference needs no exchange (BatchNorm in eval mode,
    InstanceNorm per sample); `gather=True` all-gathers the five output maps for callers that want them everywhere."""

    def __init__(self, generator):
        self.generator = generator

    def __call__(self, rgb, depth, normal, gather=False):
        world, rank = _world()
        lo, hi = shard_bounds(rgb.shape[0], world, rank)
        out = self.generator(rgb[lo:hi], depth[lo:hi], normal[lo:hi])
        if not gather or world == 1:
            return out
        sizes = [shard_bounds(rgb.shape[0], world, r) for r in range(world)]
        nmax = max(h - l for l, h in sizes)
        merged = {}
        for k, v in out.items():
            pad = v.new_zeros((nmax,) + tuple(v.shape[1:]))      # all_gather wants equal shapes: pad ragged shards
            pad[:v.shape[0]] = v
            parts = [torch.empty_like(pad) for _ in sizes]
            dist.all_gather(parts, pad)
            merged[k] = torch.cat([p[:h - l] for p, (l, h) in zip(parts, sizes)], 0)
        return merged


def allreduce_gradients(parameters, average=True):
    """ONE all-reduce over the flattened gradients of the parameters that received one (fuse_layer5 and the frozen
    NLSPN dummies never do), instead of DDP's 25 MB buckets.  Returns the number of elements reduced."""
    world, _ = _world()
    grads = [p.grad for p in parameters if p.grad is not None]
    if not grads or world == 1:
        return sum(g.numel() for g in grads)
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    if average:
        flat.div_(world)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n
    return off


def reduce_losses(losses):
    """base.py:121-132 semantics (mean over ranks of every scalar), as one all-reduce instead of one per key."""
    world, _ = _world()
    keys = sorted(losses)
    if world == 1 or not keys:
        return {k: float(losses[k]) for k in keys}
    vec = torch.stack([torch.as_tensor(losses[k], dtype=torch.float32).detach().reshape(()) for k in keys])
    dev = next((v.device for v in losses.values() if torch.is_tensor(v)), vec.device)
    vec = vec.to(dev) / world
    dist.all_reduce(vec, op=dist.ReduceOp.SUM)
    return {k: float(v) for k, v in zip(keys, vec.tolist())}
