"""Data-parallel shim for the hot path: one process per GPU, the view batch is sharded
across ranks, geometry / materials / light are replicated, and ONE all-reduce over a flat fp32 bucket of all
trainable-parameter gradients is issued per step.  The reference has no distributed code at all
(grep nccl|distributed over the tree is empty); this is the only collective the path needs:
every forward quantity is per-pixel, and the only cross-view reductions in backward are plain sums
(env-map gradient kernel.cu:203-211, texture / vertex gradients upstream of the kernel).

Per-rank losses must be the LOCAL mean over the rank's views; averaging the buckets then reproduces the
single-GPU gradient of the global mean for equal-sized shards (train.py:51-66, renderutils/ops.py:494).
The per-pixel RNG hash includes the global view index: pass `batch_offset=shard.start` to optix_env_shade.
"""
import torch
import torch.distributed as dist


def shard_views(global_batch, rank=None, world=None):
    """Contiguous, equal shards: rank r owns views [r*B/k, (r+1)*B/k).  Returns a slice."""
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    if global_batch % world != 0:
        raise ValueError("global batch %d is not divisible by world size %d (equal shards keep the mean-of-means exact)" % (global_batch, world))
    per = global_batch // world
    return slice(rank * per, (rank + 1) * per)


class GradBucket:
    """Flat parameter + gradient storage: parameters are views into `flat`, their .grad are views into
    `flat_grad`, so `all_reduce_mean()` is a single collective on one contiguous buffer (no per-tensor launches,
    no copies).  <= ~110 MB for the reference's largest configuration."""

    def __init__(self, shapes, device, dtype=torch.float32):
        self.shapes = [tuple(s) for s in shapes]
        sizes = [int(torch.Size(s).numel()) for s in self.shapes]
        self.offsets = [0]
        for n in sizes:
            self.offsets.append(self.offsets[-1] + n)
        self.flat = torch.zeros(self.offsets[-1], device=device, dtype=dtype)
        self.flat_grad = torch.zeros_like(self.flat)
        self.params = []
        for i, s in enumerate(self.shapes):
            p = self.flat[self.offsets[i]:self.offsets[i + 1]].view(s).requires_grad_(True)
            p.grad = self.flat_grad[self.offsets[i]:self.offsets[i + 1]].view(s)
            self.params.append(p)

    def zero_grad(self):
        self.flat_grad.zero_()

    def all_reduce_mean(self, group=None):
        """SUM all-reduce then divide by the world size; a no-op without an initialised process group."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM, group=group)
            self.flat_grad.div_(dist.get_world_size(group))
        return self.flat_grad

    def nbytes(self):
        return self.flat_grad.numel() * self.flat_grad.element_size()
