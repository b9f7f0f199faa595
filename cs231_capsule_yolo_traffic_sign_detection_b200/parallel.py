"""Data-parallel plumbing for the routing path: one process per GPU, batch-sharded, gradients
averaged with ONE NCCL all-reduce over a flat fp32 bucket.

The routing path has no data-path exchange: samples are independent (reference models.py:70-79
couples only capsules of the same sample), so ranks only meet in the dW / backbone-gradient
all-reduce.  `GradBucket` makes every parameter's .grad a view into one contiguous buffer, so the
backward kernels' outputs are accumulated straight into the buffer NCCL reduces.
Works with the `gloo` backend on CPU tensors too (used by the world_size-2 CPU tests)."""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """torchrun-style init (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*).  Returns (rank, world, device)."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    use_cuda = torch.cuda.is_available()
    device = torch.device('cuda', local) if use_cuda else torch.device('cpu')
    if use_cuda:
        torch.cuda.set_device(device)
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.setdefault('MASTER_PORT', '29500')
        kw = {}
        if use_cuda:
            kw['device_id'] = device
        dist.init_process_group(backend or ('nccl' if use_cuda else 'gloo'), rank=rank, world_size=world, **kw)
    return rank, world, device


def shard_bounds(n, rank, world):
    """Contiguous, balanced [lo, hi) slice of n samples for this rank (np.array_split order)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class GradBucket:
    """Flat gradient bucket: p.grad of every parameter is a view into `self.flat`."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError('no trainable parameters')
        dev, dt = self.params[0].device, self.params[0].dtype
        total = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(total, device=dev, dtype=dt)
        o = 0
        for p in self.params:
            p.grad = self.flat[o:o + p.numel()].view_as(p)
            o += p.numel()

    def zero(self):
        self.flat.zero_()

    def allreduce(self, average=True, async_op=False):
        """Sum (or average) the bucket across ranks.  Each rank computes its loss with 1/B_local
        (reference loss_fns.py:23 per rank), so averaging reproduces the single-process update on
        the concatenated batch when shards are equal."""
        if not dist.is_initialized() or dist.get_world_size() == 1:
            return None
        work = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, async_op=async_op)
        if average:
            if async_op:
                work.wait()
            self.flat.div_(dist.get_world_size())
        return work
