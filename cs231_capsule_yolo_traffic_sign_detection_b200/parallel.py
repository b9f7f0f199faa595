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
        # a rank that leaves early must not hold its peers for NCCL's default 10 minutes
        import datetime
        kw['timeout'] = datetime.timedelta(seconds=int(os.environ.get('CAPS_DIST_TIMEOUT_S', '180')))
        dist.init_process_group(backend or ('nccl' if use_cuda else 'gloo'), rank=rank, world_size=world, **kw)
    return rank, world, device


def shard_bounds(n, rank, world):
    """Contiguous, balanced [lo, hi) slice of n samples for this rank (np.array_split order)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class GradBucket:
    """Flat gradient bucket: p.grad of every parameter is a view into `self.flat`, so the backward kernels' outputs
    land in the buffer the collective reduces, and ONE all-reduce per step covers routing weights and backbone alike.

    The all-reduce can run on a side stream behind an event (`allreduce_async(ready_event)`): the routing backward
    records that event right behind the kernel that completes dW (caps_route_backward_ev), so the collective overlaps
    the rest of the backward and the next step's forward; `wait()` makes the current stream wait for it (no host block).

    `optimizer.zero_grad()` defaults to set_to_none=True since torch 2.0, which would silently detach the gradients
    from the bucket: use `bucket.zero()` instead, or call zero_grad(set_to_none=False); `allreduce*` re-binds (and
    copies) any gradient that no longer lives in the bucket, so a stale bucket is never reduced."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError('no trainable parameters')
        dev, dt = self.params[0].device, self.params[0].dtype
        total = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(total, device=dev, dtype=dt)
        self._views = []
        o = 0
        for p in self.params:
            view = self.flat[o:o + p.numel()].view_as(p)
            p.grad = view
            self._views.append(view)
            o += p.numel()
        self._work = None
        self._comm_stream = None
        self._average_pending = False

    def zero(self):
        self.flat.zero_()

    def rebind(self):
        """Puts every .grad back into the bucket (copying what was accumulated elsewhere).  Returns how many had left."""
        moved = 0
        for p, view in zip(self.params, self._views):
            g = p.grad
            if g is None:
                view.zero_()
                p.grad = view
                moved += 1
            elif g.data_ptr() != view.data_ptr():
                view.copy_(g)
                p.grad = view
                moved += 1
        return moved

    def _reduce(self, average, async_op):
        world = dist.get_world_size()
        avg_native = average and self.flat.is_cuda            # NCCL has ReduceOp.AVG; gloo does not
        op = dist.ReduceOp.AVG if avg_native else dist.ReduceOp.SUM
        work = dist.all_reduce(self.flat, op=op, async_op=async_op)
        return work, (average and not avg_native), world

    def allreduce(self, average=True, async_op=False):
        """Sum (or average) the bucket across ranks.  Each rank computes its loss with 1/B_local
        (reference loss_fns.py:23 per rank), so averaging reproduces the single-process update on
        the concatenated batch when shards are equal."""
        if not dist.is_initialized() or dist.get_world_size() == 1:
            return None
        self.rebind()
        work, need_div, world = self._reduce(average, async_op)
        if need_div:
            if async_op:
                work.wait()
            self.flat.div_(world)
        return work

    def allreduce_async(self, ready_event=None, average=True):
        """Enqueue the all-reduce on the bucket's side stream, behind `ready_event` (a torch.cuda.Event recorded when
        the last gradient of the bucket is complete; None = behind everything queued on the current stream)."""
        if not dist.is_initialized() or dist.get_world_size() == 1:
            return None
        self.rebind()
        if not self.flat.is_cuda:
            return self.allreduce(average=average)
        if self._comm_stream is None:
            self._comm_stream = torch.cuda.Stream(device=self.flat.device)
        if ready_event is None:
            ready_event = torch.cuda.Event()
            ready_event.record()
        self._comm_stream.wait_event(ready_event)
        with torch.cuda.stream(self._comm_stream):
            self._work, self._average_pending, _ = self._reduce(average, True)
        return self._work

    def wait(self):
        """The current stream waits for the pending asynchronous all-reduce (stream-level: the host does not block)."""
        if self._work is not None:
            self._work.wait()
            if self._average_pending:
                self.flat.div_(dist.get_world_size())
            self._work = None
            self._average_pending = False
