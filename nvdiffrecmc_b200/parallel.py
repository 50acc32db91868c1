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
    no copies).  <= ~110 MB for the reference's largest configuration.

    `GradBucket(shapes, device)` creates fresh leaf parameters; `GradBucket.adopt(params)` re-points EXISTING
    parameters (the nn.Parameters the reference's modules own: light.base, the material textures, v_pos / sdf / deform,
    train.py:340-356) into the flat storage in place, so the modules and the optimizer keep the very same tensor objects.

    The aliasing `p.grad is a view of flat_grad` can be broken from outside: torch's `optimizer.zero_grad()` defaults to
    set_to_none=True (the reference calls it every iteration, train.py:407-411), after which autograd allocates fresh
    .grad tensors.  `all_reduce_mean()` therefore re-establishes the aliasing first (`sync_views`): a foreign .grad is
    copied into its slice, a missing one zeroes its slice, and .grad is pointed back at the bucket -- never a stale reduce."""

    def __init__(self, shapes, device, dtype=torch.float32, _params=None):
        self.shapes = [tuple(s) for s in shapes]
        sizes = [int(torch.Size(s).numel()) for s in self.shapes]
        self.offsets = [0]
        for n in sizes:
            self.offsets.append(self.offsets[-1] + n)
        self.flat = torch.zeros(self.offsets[-1], device=device, dtype=dtype)
        self.flat_grad = torch.zeros_like(self.flat)
        self.params = []
        for i, s in enumerate(self.shapes):
            if _params is None:
                p = self.flat[self.offsets[i]:self.offsets[i + 1]].view(s).requires_grad_(True)
            else:
                p = _params[i]
                with torch.no_grad():
                    self.flat[self.offsets[i]:self.offsets[i + 1]].copy_(p.detach().reshape(-1))
                    p.data = self.flat[self.offsets[i]:self.offsets[i + 1]].view(s)
            p.grad = self._grad_view(i)
            self.params.append(p)

    @classmethod
    def adopt(cls, params):
        """Bucket over existing leaf tensors / nn.Parameters (all on one device, one dtype): their storage moves into `flat`."""
        params = [p for p in params if p.requires_grad]
        if not params:
            raise ValueError("GradBucket.adopt: no trainable parameters")
        dev, dt = params[0].device, params[0].dtype
        for p in params:
            if p.device != dev or p.dtype != dt:
                raise ValueError("GradBucket.adopt: parameters must share device and dtype (got %s/%s and %s/%s)" % (dev, dt, p.device, p.dtype))
        return cls([p.shape for p in params], dev, dt, _params=params)

    def _grad_view(self, i):
        return self.flat_grad[self.offsets[i]:self.offsets[i + 1]].view(self.shapes[i])

    def zero_grad(self):
        """Zero the bucket in one launch and keep every .grad aliased to it."""
        self.flat_grad.zero_()
        for i, p in enumerate(self.params):
            g = p.grad
            if g is None or g.data_ptr() != self.flat_grad.data_ptr() + self.offsets[i] * self.flat_grad.element_size():
                p.grad = self._grad_view(i)

    def sync_views(self):
        """Re-establish `p.grad aliases flat_grad` after something replaced or dropped the .grad tensors."""
        esz = self.flat_grad.element_size()
        base = self.flat_grad.data_ptr()
        for i, p in enumerate(self.params):
            g = p.grad
            if g is not None and g.data_ptr() == base + self.offsets[i] * esz:
                continue
            view = self._grad_view(i)
            with torch.no_grad():
                if g is None:
                    view.zero_()
                else:
                    view.copy_(g)
            p.grad = view

    def all_reduce_mean(self, group=None):
        """SUM all-reduce then divide by the world size; a no-op (besides `sync_views`) without an initialised process group."""
        self.sync_views()
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM, group=group)
            self.flat_grad.div_(dist.get_world_size(group))
        return self.flat_grad

    def nbytes(self):
        return self.flat_grad.numel() * self.flat_grad.element_size()


# ------------------------------------------------------------------------------------------------
# Training-loop shim: what train.py:313-494 needs around the kernels to run one process per GPU
# without editing the reference's files -- rank-consistent randomness, sharded sampling of the view batch, and the all-reduce
# hooked in front of optimizer.step().
# ------------------------------------------------------------------------------------------------
def sync_seed(seed=None, group=None, device=None):
    """One integer agreed on by all ranks (rank 0's value wins).  The reference draws its per-call seeds from the process-global
    numpy RNG (render/optixutils/ops.py:83,100) and its per-iteration `rnd_seed` from a module global (render/render.py:19); ranks
    must use the same stream or the shards stop being shards of ONE Monte-Carlo estimate."""
    import numpy as np
    if seed is None:
        seed = int(np.random.randint(2 ** 31))
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        t = torch.tensor([int(seed)], dtype=torch.int64, device=device if device is not None else "cpu")
        dist.broadcast(t, src=0, group=group)
        seed = int(t.item())
    return int(seed)


class StepRNG:
    """Rank-consistent randomness for the host-side draws of a training step (camera jitter `torch.normal`, render/render.py:50,63;
    random background, train.py:94): `for_step(it)` returns a generator that is identical on every rank for iteration `it`;
    `for_view(it, global_view)` one per global view index, so a rank that owns views [a, b) draws exactly what a single process
    would have drawn for those views."""

    def __init__(self, seed, device="cpu"):
        self.seed, self.device = int(seed), device

    def _gen(self, *keys):
        h = self.seed & 0xFFFFFFFF
        for k in keys:
            h = (h * 747796405 + 2891336453 + int(k)) & 0xFFFFFFFFFFFFFFFF      # same LCG constants as the kernel's PCG (kernel.cu:33)
            h ^= h >> 29
        return torch.Generator(device=self.device).manual_seed(h & 0x7FFFFFFFFFFFFFFF)

    def for_step(self, it):
        return self._gen(1, it)

    def for_view(self, it, global_view):
        return self._gen(2, it, global_view)


def shard_indices(n_items, global_batch, it, seed, rank=None, world=None):
    """Dataset indices of THIS rank for iteration `it`: every rank computes the same epoch permutation (seeded by `seed` and the epoch
    number), the global batch of iteration `it` is a contiguous window of it, and rank r takes `shard_views` of that window.  The union
    over ranks is exactly the batch a single process with the same seed would load (dataset sampling of train.py:355-360 made
    rank-consistent)."""
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    per_epoch = max(n_items // global_batch, 1)
    epoch, k = divmod(int(it), per_epoch)
    g = torch.Generator().manual_seed((int(seed) * 1000003 + epoch) & 0x7FFFFFFFFFFFFFFF)
    perm = torch.randperm(n_items, generator=g)
    window = perm[(k * global_batch) % n_items:(k * global_batch) % n_items + global_batch]
    if window.numel() < global_batch:                                   # dataset smaller than one batch: wrap around
        window = torch.cat([window, perm[:global_batch - window.numel()]])
    return window[shard_views(global_batch, rank, world)]


def hook_optimizer(optimizer, bucket, group=None):
    """Make `optimizer.step()` average the gradient bucket over the ranks first (ONE collective), so the reference's loop body
    `optimizer.zero_grad(); ...; total_loss.backward(); ...; optimizer.step(); scheduler.step()` (train.py:407-461) needs no edit.

    * the all-reduce is a torch step PRE-HOOK (`Optimizer.register_step_pre_hook`), so `optimizer.step` stays the bound method that
      `torch.optim.lr_scheduler.LambdaLR` wraps / inspects (train.py:349,353,356 build one scheduler per optimizer) -- hooking
      before or after the scheduler is constructed both work;
    * `optimizer.zero_grad()` is routed to `bucket.zero_grad()` (one memset, .grad stays aliased) regardless of `set_to_none`;
      even without that, `all_reduce_mean` re-aliases foreign / missing .grad tensors (`GradBucket.sync_views`).
    Returns the optimizer."""
    import types
    optimizer.register_step_pre_hook(lambda opt, args, kwargs: (bucket.all_reduce_mean(group), None)[1])

    def zero_grad(self, set_to_none=True):
        bucket.zero_grad()

    optimizer.zero_grad = types.MethodType(zero_grad, optimizer)
    return optimizer
