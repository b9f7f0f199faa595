"""Optional sync-free side runner for the reference's epoch loops.

The reference's `main.train` / `main.evaluate` (main.py:38-95, :98-140) stall the host three times per step:
`y_hat_bch.data.cpu().numpy()` (main.py:68) and two `loss.item()` (main.py:74, :77).  Once the routing layer takes
tens of microseconds instead of hundreds of milliseconds these stalls, and the synchronous `.to(device)` of every
batch (main.py:57-59), are what a step costs.  north_star keeps `main.py` unchanged, so this is a SIDE runner with
the same signature, the same batching (`np.array_split`, main.py:43-44), the same step order and the same return
values, that a caller may use instead of `main.train` / `main.evaluate`:

  * every batch is staged in pinned host memory and copied on a copy stream one step ahead of the compute;
  * predictions go to ONE preallocated pinned host array with asynchronous copies (no per-step `.cpu()`);
  * per-step losses stay in a device vector; the host reads it once per epoch and then accumulates exactly like
    main.py:77 (`avg_loss += loss / n_batch`, in batch order, in double precision): bit-identical `avg_loss`;
  * one synchronisation per epoch;
  * `graph=True` (CUDA only): the whole step -- forward, loss, `zero_grad`, backward, optimizer step -- is captured in
    ONE CUDA graph per batch size and replayed (at the reference's batch sizes a step is ~150 launches of a few
    microseconds each and the host cannot issue them as fast as the GPU retires them).  The first two batches of a
    size run eagerly (they are real steps and warm up cuDNN / the kernels' attributes); the optimizer must be
    built with `capturable=True`.  The reference's `capsule_loss` creates `torch.eye(n)` on the host and copies it
    (loss_fns.py:14-15); under capture the step runs inside `torch.device(device)`, which makes that factory call
    allocate on the device, so the unchanged loss function is capturable.

On a CPU device the runner degrades to the reference's loop order without streams (used by the CPU tests)."""
import weakref

import numpy as np
import torch

__all__ = ['train', 'evaluate', 'reset_graphs']

MAX_METRIC_SAMPLES = 1000      # config.py:53 (config.max_metric_samples)


def _to_model_input(x_np):
    """main.py:57: `torch.from_numpy(x).float().permute(0, 3, 1, 2)`; made NCHW-contiguous on the host so that the
    staged copy is one flat transfer (and so that models.py:81's `.view` sees the layout torch 0.4 produced)."""
    return torch.from_numpy(np.ascontiguousarray(x_np)).float().permute(0, 3, 1, 2).contiguous()


class _Stager:
    """Pinned staging of (x, y) batches; the copies run on their own stream one batch ahead of the compute.  The pinned
    buffers are a ring allocated once per epoch (cudaHostAlloc is slow and synchronises): a slot is rewritten only after
    the event behind its previous host-to-device copy has completed.  The device tensors of a batch come fresh from
    torch's caching allocator, which is told (`record_stream`) that the compute stream reads them."""
    RING = 4

    def __init__(self, x_split, y_split, device):
        self.x_split, self.y_split, self.device = x_split, y_split, device
        self.cuda = device.type == 'cuda'
        self.slots = {}
        if self.cuda:
            self.copy_stream = torch.cuda.Stream(device)
            nmax = max(len(v) for v in y_split)
            x0, y0 = x_split[0], y_split[0]
            self.x_pin = [torch.empty((nmax, x0.shape[3], x0.shape[1], x0.shape[2]), dtype=torch.float32, pin_memory=True) for _ in range(self.RING)]
            self.y_pin = [torch.empty((nmax,) + tuple(y0.shape[1:]), dtype=torch.from_numpy(y0[:0]).dtype, pin_memory=True) for _ in range(self.RING)]
            self.copied = [None] * self.RING

    def submit(self, i):
        if i >= len(self.y_split):
            return
        if not self.cuda:
            x_h, y_h = _to_model_input(self.x_split[i]), torch.from_numpy(np.ascontiguousarray(self.y_split[i]))
            self.slots[i] = (x_h.to(self.device), y_h.to(self.device), None)
            return
        r, n = i % self.RING, len(self.y_split[i])
        if self.copied[r] is not None:
            self.copied[r].synchronize()        # the copy issued RING batches ago: long complete, no stall in steady state
        x_p, y_p = self.x_pin[r][:n], self.y_pin[r][:n]
        x_p.copy_(torch.from_numpy(self.x_split[i]).permute(0, 3, 1, 2))      # main.py:57: .float().permute(0, 3, 1, 2), NCHW-contiguous
        y_p.copy_(torch.from_numpy(self.y_split[i]))
        with torch.cuda.stream(self.copy_stream):
            x_d = x_p.to(self.device, non_blocking=True)
            y_d = y_p.to(self.device, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        self.copied[r] = ev
        self.slots[i] = (x_d, y_d, ev)

    def take(self, i):
        x_d, y_d, ev = self.slots.pop(i)
        if ev is not None:
            cur = torch.cuda.current_stream(self.device)
            cur.wait_event(ev)
            x_d.record_stream(cur)
            y_d.record_stream(cur)
        return x_d, y_d


class _GraphedStep:
    """One step (forward + loss [+ zero_grad + backward + optimizer.step]) captured over static input buffers."""

    def __init__(self, model, optimizer, loss_fn, params, use_recon, training, x0, y0):
        dev = x0.device
        self.x, self.y = torch.empty_like(x0), torch.empty_like(y0)
        self.x.copy_(x0)
        self.y.copy_(y0)
        self.graph = torch.cuda.CUDAGraph()
        torch.cuda.synchronize(dev)
        with torch.cuda.graph(self.graph), torch.device(dev):
            if use_recon:
                y_hat, recon = model(self.x, self.y, True)
                loss = loss_fn(y_hat, self.y, params, self.x, recon)
            else:
                y_hat = model(self.x)
                loss = loss_fn(y_hat, self.y, params)
            if training:
                optimizer.zero_grad(set_to_none=True)
                loss.backward()
                optimizer.step()
            self.y_hat, self.loss = y_hat.detach(), loss.detach()

    def run(self, x_bch, y_bch):
        self.x.copy_(x_bch)
        self.y.copy_(y_bch)
        self.graph.replay()
        return self.y_hat, self.loss


def _eager_step(model, optimizer, loss_fn, params, use_recon, training, x_bch, y_bch):
    """main.py:61-72.  A function of its own so that nothing keeps the step's autograd graph alive after it returns
    (a live AccumulateGrad node of an eager step, bound to the default stream, would invalidate a later capture)."""
    if use_recon:
        y_hat_bch, recon = model(x_bch, y_bch, True)                              # main.py:62
        loss = loss_fn(y_hat_bch, y_bch, params, x_bch, recon)                    # main.py:63
    else:
        y_hat_bch = model(x_bch)                                                  # main.py:65
        loss = loss_fn(y_hat_bch, y_bch, params)                                  # main.py:66
    if training:
        optimizer.zero_grad()                                                     # main.py:70-72
        loss.backward()
        optimizer.step()
    return y_hat_bch.detach(), loss.detach()


_EAGER_STEPS_BEFORE_CAPTURE = 2


def _graph_cache(model, optimizer, training):
    """Graphs live as long as the model and belong to ONE optimizer object: {(batch size, training): _GraphedStep or the
    number of eager steps taken so far}.  A different optimizer (or none after one) starts over -- a captured step
    updates the state tensors of the optimizer it was captured with.  Parameters must stay where they are (in-place
    `load_state_dict` is fine; `model.to(...)` / re-created parameters need `reset_graphs(model)`)."""
    if training and not all(g.get('capturable', False) for g in optimizer.param_groups):
        raise ValueError('runner: graph=True needs an optimizer built with capturable=True (its step counter must live on the device)')
    slot = model.__dict__.get('_caps_runner_graphs')
    owner = slot['optimizer']() if slot is not None and slot['optimizer'] is not None else None
    if slot is None or (training and owner is not optimizer):
        slot = {'optimizer': weakref.ref(optimizer) if optimizer is not None else (slot['optimizer'] if slot else None), 'steps': {}}
        model.__dict__['_caps_runner_graphs'] = slot
    return slot['steps']


def reset_graphs(model):
    """Drops the CUDA graphs `train` / `evaluate(graph=True)` captured for this model."""
    model.__dict__.pop('_caps_runner_graphs', None)


def _epoch(x, y, model, optimizer, loss_fn, params, training, graph=False):
    device = torch.device(params.device)
    total = len(y)
    n_batch = (total + params.batch_size - 1) // params.batch_size
    x_split, y_split = np.array_split(x, n_batch), np.array_split(y, n_batch)
    cuda = device.type == 'cuda'
    use_recon = params.model == 'capsule' and params.recon        # main.py:61
    stager = _Stager(x_split, y_split, device)
    stager.submit(0)
    losses = torch.zeros(n_batch, device=device)
    y_hat_host, offset = None, 0
    cache = _graph_cache(model, optimizer, training) if (graph and cuda) else None
    for i in range(n_batch):
        stager.submit(i + 1)
        x_bch, y_bch = stager.take(i)
        key = (x_bch.shape[0], training)
        step = cache.get(key, 0) if cache is not None else 0
        if cache is not None and step == _EAGER_STEPS_BEFORE_CAPTURE:
            step = cache[key] = _GraphedStep(model, optimizer, loss_fn, params, use_recon, training, x_bch, y_bch)
        if isinstance(step, _GraphedStep):
            yh, loss = step.run(x_bch, y_bch)
        else:
            yh, loss = _eager_step(model, optimizer, loss_fn, params, use_recon, training, x_bch, y_bch)
            if cache is not None:
                cache[key] = step + 1
        if y_hat_host is None:
            y_hat_host = torch.empty((total,) + tuple(yh.shape[1:]), dtype=yh.dtype, pin_memory=cuda)
        y_hat_host[offset:offset + yh.shape[0]].copy_(yh, non_blocking=True)      # main.py:68 without the stall
        offset += yh.shape[0]
        losses[i] = loss
    if cuda:
        torch.cuda.synchronize(device)                                            # the epoch's only synchronisation
    avg_loss = 0
    for v in losses.cpu().tolist():                                               # main.py:77, same order, same arithmetic
        avg_loss += v / n_batch
    return avg_loss, y_hat_host.numpy()


def _metric(y, y_hat, metric, params, if_eval, no_metric):
    metric_score = -1
    if if_eval and not no_metric:                                                 # main.py:85-90
        n = y.shape[0]
        if n > MAX_METRIC_SAMPLES:
            i = np.random.choice(n, MAX_METRIC_SAMPLES).astype(int)
            y, y_hat = y[i], y_hat[i]
        metric_score = metric(y, y_hat, params)
    return metric_score


def train(x, y, model, optimizer, loss_fn, metric, params, if_eval=True, no_metric=False, shuffle=True, graph=False):
    """main.train (main.py:38-95) without per-step host synchronisation.  `no_metric` is main.py's `args.no_metric`;
    `shuffle=False` skips `utils.shuffle` (main.py:41) for reproducible comparisons; `graph=True` replays each step
    from a CUDA graph (module docstring).  -> (avg_loss, metric_score)."""
    model.train()
    if shuffle:
        i = np.random.permutation(len(y))                                         # utils.py:146-148
        x, y = x[i], y[i]
    avg_loss, y_hat = _epoch(x, y, model, optimizer, loss_fn, params, training=True, graph=graph)
    return avg_loss, _metric(y, y_hat, metric, params, if_eval, no_metric)


def evaluate(x, y, model, loss_fn, metric, params, if_eval=True, no_metric=False, graph=False):
    """main.evaluate (main.py:98-140) without per-step host synchronisation.  -> (avg_loss, metric_score)."""
    model.eval()
    with torch.no_grad():
        avg_loss, y_hat = _epoch(x, y, model, None, loss_fn, params, training=False, graph=graph)
    return avg_loss, _metric(y, y_hat, metric, params, if_eval, no_metric)
