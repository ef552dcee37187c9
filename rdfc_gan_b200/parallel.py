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
        if hi == lo:
            raise RuntimeError(f"rank {rank} of {world} owns no image of a batch of {rgb.shape[0]}: use at most one rank per image")
        out = self.generator(rgb[lo:hi], depth[lo:hi], normal[lo:hi])
        if isinstance(out, tuple):              # DCVGANGenerator-style 5-tuple
            out = dict(zip(('depth_map_1', 'confidence_map_1', 'depth_map_2', 'confidence_map_2', 'pred_depth'), out))
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


class GradientBucket:
    """ONE pre-flattened gradient buffer per dtype for a FIXED parameter list: every ``p.grad`` is a view into it, so the
    all-reduce needs no ``torch.cat`` and no copy-back, and every rank reduces the same number of elements whatever subset of
    parameters received a gradient this step (parameters that got none contribute zeros -- DDP's find_unused_parameters
    behaviour; fuse_layer5 and the frozen NLSPN dummies never get one).  Replaces DistributedDataParallel's
    25 MB buckets (C/lib/models/rdfc_gan.py:102-119)."""

    def __init__(self, parameters):
        self.params = [p for p in parameters if p.requires_grad]
        self.flat = {}
        by_dtype = {}
        for p in self.params:
            by_dtype.setdefault((p.dtype, p.device), []).append(p)
        for key, ps in by_dtype.items():
            buf = torch.zeros(sum(p.numel() for p in ps), dtype=key[0], device=key[1])
            off = 0
            for p in ps:
                p.grad = buf[off:off + p.numel()].view_as(p)
                off += p.numel()
            self.flat[key] = buf

    def numel(self):
        return sum(b.numel() for b in self.flat.values())

    def zero(self):
        """Instead of optimizer.zero_grad(set_to_none=True), which would detach the views."""
        for b in self.flat.values():
            b.zero_()

    def _rebind(self):
        # autograd may have replaced a .grad (e.g. after zero_grad(set_to_none=True)): copy it back into its slot
        for (dtype, dev), buf in self.flat.items():
            off = 0
            for p in (q for q in self.params if (q.dtype, q.device) == (dtype, dev)):
                view = buf[off:off + p.numel()].view_as(p)
                if p.grad is None:
                    view.zero_()
                    p.grad = view
                elif p.grad.data_ptr() != view.data_ptr():
                    view.copy_(p.grad)
                    p.grad = view
                off += p.numel()

    def allreduce(self, average=True):
        """Sum (then / world) over the ranks, in place.  Returns the number of elements reduced."""
        world, _ = _world()
        self._rebind()
        if world > 1:
            for buf in self.flat.values():
                dist.all_reduce(buf, op=dist.ReduceOp.SUM)
                if average:
                    buf.div_(world)
        return self.numel()


def allreduce_gradients(parameters, average=True):
    """One-shot form of GradientBucket for callers that do not keep a bucket: reduces over the FIXED list of parameters that
    require grad (missing gradients count as zeros, so all ranks agree on the size), one all-reduce per dtype."""
    world, _ = _world()
    params = [p for p in parameters if p.requires_grad]
    groups = {}
    for p in params:
        groups.setdefault((p.dtype, p.device), []).append(p)
    total = 0
    for ps in groups.values():
        total += sum(p.numel() for p in ps)
        if world == 1:
            continue
        flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in ps])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        if average:
            flat.div_(world)
        off = 0
        for p in ps:
            n = p.numel()
            if p.grad is None:
                p.grad = flat[off:off + n].view_as(p).clone()
            else:
                p.grad.copy_(flat[off:off + n].view_as(p))
            off += n
    return total


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
