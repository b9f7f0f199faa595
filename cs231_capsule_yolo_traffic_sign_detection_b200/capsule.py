"""Host-side mirror of the reference's capsule layer, over the C ABI in include/caps_routing.h.

`CapsuleLayer` keeps the reference's constructor, attributes, parameter names/shapes and output
shape (reference models.py:46-83), so it drops into `models.CapsuleNet` / `models.DarkCapsuleNet`
(assign `models.CapsuleLayer = CapsuleLayer` before building the model) with `main.py`,
`loss_fns.py`, `predict_fns.py` unchanged.  The caps->caps branch runs in the hand-written
sm_100a kernels.  The conv->caps branch (n_nodes == -1, the step directly before the routing
layer) runs its n_caps convolutions as ONE cuDNN convolution over the concatenated weights and
does the views + cat + squash in one kernel (caps_primary_squash); on CPU tensors it is the
reference's stock PyTorch code.

PyTorch is plumbing here (device memory, streams, autograd graph); the arithmetic is in
libcaps_routing.so.  There is no CPU fallback: CPU tensors on the routing branch raise.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _cabi


def _ptr(t):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _check_routing_inputs(u, W5):
    if not (u.is_cuda and W5.is_cuda):
        raise RuntimeError('capsule routing runs on a CUDA device only (no CPU fallback): '
                           'got u on %s, route_weights on %s' % (u.device, W5.device))
    if u.dtype != torch.float32 or W5.dtype != torch.float32:
        raise RuntimeError('capsule routing is fp32 only: got %s / %s' % (u.dtype, W5.dtype))
    if u.dim() != 3 or W5.dim() != 5 or W5.shape[0] != 1:
        raise RuntimeError('expected u [B,N,K] and route_weights [1,N,C,K,D], got %s and %s'
                           % (tuple(u.shape), tuple(W5.shape)))
    if u.shape[1] != W5.shape[1] or u.shape[2] != W5.shape[3]:
        raise RuntimeError('shape mismatch: u %s vs route_weights %s' % (tuple(u.shape), tuple(W5.shape)))


def _forward_impl(u, W5, n_iter, with_grad, want_c):
    L = _cabi.lib()
    B, N, K = u.shape
    _, _, C, _, D = W5.shape
    u = u.contiguous()
    W = W5.contiguous()
    v = torch.empty((B, C, D), device=u.device, dtype=torch.float32)
    c = torch.empty((B, N, C), device=u.device, dtype=torch.float32) if want_c else None
    nbytes = L.caps_route_workspace_bytes(B, N, C, K, D, n_iter, int(with_grad))
    if nbytes == 0:
        raise RuntimeError('capsule routing: unsupported dims B=%d N=%d C=%d K=%d D=%d n_iter=%d '
                           '(K must be 8, D <= 48, n_iter <= 5)' % (B, N, C, K, D, n_iter))
    ws = torch.empty((nbytes,), device=u.device, dtype=torch.uint8)
    with torch.cuda.device(u.device):
        _cabi.check(L.caps_route_forward(_ptr(u), _ptr(W), _ptr(v), _ptr(c), _ptr(ws), nbytes,
                                         B, N, C, K, D, n_iter, int(with_grad), _stream()),
                    'caps_route_forward')
    return u, W, v, c, ws


def _backward_impl(u, W, ws, grad_v, y, margin_scale, loss_grad, n_iter, need_du):
    L = _cabi.lib()
    B, N, K = u.shape
    _, _, C, _, D = W.shape
    dW = torch.empty_like(W)
    du = torch.empty_like(u) if need_du else None
    if grad_v is not None:
        grad_v = grad_v.contiguous()
    with torch.cuda.device(u.device):
        _cabi.check(L.caps_route_backward(_ptr(u), _ptr(W), _ptr(grad_v), _ptr(y), float(margin_scale),
                                          _ptr(loss_grad), _ptr(du), _ptr(dW), _ptr(ws), ws.numel(),
                                          B, N, C, K, D, n_iter, _stream()),
                    'caps_route_backward')
    return du, dW


class _RoutingFn(torch.autograd.Function):
    """v = routing(u, W): autograd node over caps_route_forward / caps_route_backward."""

    @staticmethod
    def forward(ctx, u, W5, n_iter, want_c):
        _check_routing_inputs(u, W5)
        ctx.set_materialize_grads(False)
        with_grad = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        u_c, W_c, v, c, ws = _forward_impl(u, W5, n_iter, with_grad, want_c)
        if with_grad:
            ctx.save_for_backward(u_c, W_c)
            ctx.ws = ws
            ctx.n_iter = n_iter
        if want_c:
            ctx.mark_non_differentiable(c)
            return v, c
        return v

    @staticmethod
    def backward(ctx, grad_v, *unused):
        u, W = ctx.saved_tensors
        du, dW = _backward_impl(u, W, ctx.ws, grad_v, None, 0.0, None, ctx.n_iter, ctx.needs_input_grad[0])
        return du, (dW if ctx.needs_input_grad[1] else None), None, None


class _RoutingMarginLossFn(torch.autograd.Function):
    """(v, loss) = routing + margin loss; the backward kernel adds the margin-loss gradient itself
    (reference loss_fns.py:12-17,23 on scores = |v|, models.py:117), so only the EXTRA gradient
    flowing into v (decoder / coordinate losses) comes in through autograd."""

    @staticmethod
    def forward(ctx, u, W5, y, n_iter):
        _check_routing_inputs(u, W5)
        ctx.set_materialize_grads(False)
        L = _cabi.lib()
        with_grad = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        u_c, W_c, v, _, ws = _forward_impl(u, W5, n_iter, with_grad, False)
        B, C, D = v.shape
        y = y.to(device=u.device, dtype=torch.int64).contiguous()
        loss = torch.empty((), device=u.device, dtype=torch.float32)
        scratch = torch.empty((_cabi.MARGIN_SCRATCH_FLOATS,), device=u.device, dtype=torch.float32)
        with torch.cuda.device(u.device):
            _cabi.check(L.caps_margin_loss(_ptr(v), _ptr(y), 1.0 / B, _ptr(loss), None, _ptr(scratch), B, C, D, _stream()),
                        'caps_margin_loss')
        if with_grad:
            ctx.save_for_backward(u_c, W_c, y)
            ctx.ws = ws
            ctx.n_iter = n_iter
        return v, loss

    @staticmethod
    def backward(ctx, grad_v, grad_loss):
        u, W, y = ctx.saved_tensors
        B = u.shape[0]
        if grad_loss is None:
            y_arg, lg = None, None
        else:
            y_arg, lg = y, grad_loss.to(torch.float32).contiguous()
        du, dW = _backward_impl(u, W, ctx.ws, grad_v, y_arg, 1.0 / B, lg, ctx.n_iter, ctx.needs_input_grad[0])
        return du, (dW if ctx.needs_input_grad[1] else None), None, None


class _SquashFn(torch.autograd.Function):
    """squash over the last dim (reference models.py:64-67) through caps_squash."""

    @staticmethod
    def forward(ctx, x):
        L = _cabi.lib()
        xc = x.contiguous()
        y = torch.empty_like(xc)
        D = xc.shape[-1]
        rows = xc.numel() // D
        with torch.cuda.device(x.device):
            _cabi.check(L.caps_squash(_ptr(xc), _ptr(y), rows, D, _stream()), 'caps_squash')
        ctx.save_for_backward(xc)
        return y

    @staticmethod
    def backward(ctx, gy):
        (xc,) = ctx.saved_tensors
        L = _cabi.lib()
        gy = gy.contiguous()
        gx = torch.empty_like(xc)
        D = xc.shape[-1]
        with torch.cuda.device(xc.device):
            _cabi.check(L.caps_squash_backward(_ptr(xc), _ptr(gy), _ptr(gx), xc.numel() // D, D, _stream()),
                        'caps_squash_backward')
        return gx


class _PrimarySquashFn(torch.autograd.Function):
    """u [B, Cc*H*W, K] = squash_k(conv [B, K*Cc, H, W]): the primary-capsule tail (reference
    models.py:81-82) through caps_primary_squash / caps_primary_squash_backward."""

    @staticmethod
    def forward(ctx, conv, n_caps):
        L = _cabi.lib()
        conv = conv.contiguous()
        B, KC, H, W = conv.shape
        Cc, HW = KC // n_caps, H * W
        u = torch.empty((B, Cc * HW, n_caps), device=conv.device, dtype=torch.float32)
        with torch.cuda.device(conv.device):
            _cabi.check(L.caps_primary_squash(_ptr(conv), _ptr(u), B, n_caps, Cc, HW, _stream()), 'caps_primary_squash')
        ctx.save_for_backward(conv)
        ctx.n_caps = n_caps
        return u

    @staticmethod
    def backward(ctx, du):
        (conv,) = ctx.saved_tensors
        L = _cabi.lib()
        B, KC, H, W = conv.shape
        K = ctx.n_caps
        dconv = torch.empty_like(conv)
        with torch.cuda.device(conv.device):
            _cabi.check(L.caps_primary_squash_backward(_ptr(conv), _ptr(du.contiguous()), _ptr(dconv), B, K, KC // K, H * W,
                                                       _stream()), 'caps_primary_squash_backward')
        return dconv, None


def primary_capsules(x, convs):
    """x [B,in_C,H,W] through the K capsule convolutions `convs` (an nn.ModuleList of identical-shape
    Conv2d) -> squashed u [B, out_C*H'*W', K]: one convolution over the concatenated weights (the
    parameters stay K separate tensors, so state_dict names/shapes are the reference's), then one
    kernel for the views + cat + squash."""
    c0 = convs[0]
    w = torch.cat([m.weight for m in convs], dim=0)
    b = None if c0.bias is None else torch.cat([m.bias for m in convs], dim=0)
    conv = F.conv2d(x, w, b, stride=c0.stride, padding=c0.padding, dilation=c0.dilation, groups=c0.groups)
    return _PrimarySquashFn.apply(conv, len(convs))


class _DarkRegroupFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, G):
        L = _cabi.lib()
        x = x.contiguous()
        B, Cch = x.shape[0], x.shape[1]
        u = torch.empty((G * B, 2 * Cch, 8), device=x.device, dtype=torch.float32)
        with torch.cuda.device(x.device):
            _cabi.check(L.caps_dark_regroup(_ptr(x), _ptr(u), B, Cch, G, _stream()), 'caps_dark_regroup')
        ctx.shape, ctx.G = x.shape, G
        return u

    @staticmethod
    def backward(ctx, du):
        L = _cabi.lib()
        dx = torch.empty(ctx.shape, device=du.device, dtype=torch.float32)
        with torch.cuda.device(du.device):
            _cabi.check(L.caps_dark_regroup_backward(_ptr(du.contiguous()), _ptr(dx), ctx.shape[0], ctx.shape[1], ctx.G, _stream()),
                        'caps_dark_regroup_backward')
        return dx, None


def dark_regroup(x, n_grid):
    """DarkCapsuleNet's cell regroup (reference models.py:393-399) in one kernel: feature map
    x [B, Cch, H, W] with H*W == 16*n_grid**2 -> u [n_grid**2 * B, 2*Cch, 8], the routing layer's
    input (cell-major batch: row q*B + b).  Same values and order as the reference's
    view / chunk / permute / contiguous / cat sequence."""
    G = int(n_grid) ** 2
    if x.dim() != 4 or x.shape[2] * x.shape[3] != 16 * G or x.shape[1] % 8:
        raise RuntimeError('dark_regroup: expected [B, Cch (multiple of 8), H, W] with H*W == 16*n_grid^2, got %s'
                           % (tuple(x.shape),))
    if not x.is_cuda or x.dtype != torch.float32:
        raise RuntimeError('dark_regroup runs on CUDA fp32 tensors only')
    return _DarkRegroupFn.apply(x, G)


class _DarkLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, v, y):
        L = _cabi.lib()
        B, G, Y = y.shape[0], y.shape[1] * y.shape[2], y.shape[3]
        v2 = v.reshape(G * B, 5).contiguous()
        y = y.to(device=v.device, dtype=torch.float32).contiguous()
        loss = torch.empty((), device=v.device, dtype=torch.float32)
        grad = torch.empty_like(v2) if ctx.needs_input_grad[0] else None
        scratch = torch.empty((_cabi.MARGIN_SCRATCH_FLOATS,), device=v.device, dtype=torch.float32)
        with torch.cuda.device(v.device):
            _cabi.check(L.caps_dark_loss(_ptr(v2), _ptr(y), 1.0 / B, _ptr(loss), _ptr(grad), _ptr(scratch), B, G, Y, _stream()),
                        'caps_dark_loss')
        ctx.save_for_backward(grad)
        ctx.vshape = v.shape
        return loss

    @staticmethod
    def backward(ctx, gl):
        (grad,) = ctx.saved_tensors
        return (grad * gl).reshape(ctx.vshape), None


def dark_capsule_loss(v, y):
    """Reference `darkcapsule_loss(caps, y, params)` (loss_fns.py:187-204, recon off) taken directly on
    the routing layer's output: v [g*g*B, ..., 5] (any shape with g*g*B*5 elements in the routing
    batch order q*B + b, e.g. the layer's [g*g*B,1,1,1,5]) and the label tensor y [B,g,g,>=5].
    Loss value and its gradient w.r.t. v come out of one kernel."""
    if not v.is_cuda or v.dtype != torch.float32:
        raise RuntimeError('dark_capsule_loss runs on CUDA fp32 tensors only')
    if y.dim() != 4 or y.shape[3] < 5 or v.numel() != y.shape[0] * y.shape[1] * y.shape[2] * 5:
        raise RuntimeError('dark_capsule_loss: expected v with B*g*g*5 elements and y [B,g,g,>=5], got %s and %s'
                           % (tuple(v.shape), tuple(y.shape)))
    return _DarkLossFn.apply(v, y)


def dynamic_routing(u, route_weights, n_iter=3, return_couplings=False):
    """u [B,N,K], route_weights [1,N,C,K,D]  ->  v [B,C,D] (and the last couplings c [B,N,C])."""
    return _RoutingFn.apply(u, route_weights, int(n_iter), bool(return_couplings))


def routing_margin_loss(u, route_weights, y, n_iter=3):
    """Fused path: returns (v [B,C,D], margin loss) with the margin-loss gradient computed inside the
    routing backward kernel.  loss == reference capsule_loss(|v|, y) with recon off."""
    return _RoutingMarginLossFn.apply(u, route_weights, y, int(n_iter))


class CapsuleLayer(nn.Module):
    """Drop-in for reference models.CapsuleLayer (models.py:46-83): same signature, attributes,
    parameter names and shapes, same seeded initialisation draw, same output shapes."""

    def __init__(self, params, n_caps, n_nodes, in_C, out_C, kernel=None, stride=None, n_iter=3):
        super(CapsuleLayer, self).__init__()
        self.params = params
        # the reference hard-codes n_iter at its call sites (models.py:48, :93, :368); an optional `n_iter` entry in
        # params.json (utils.Params attribute) overrides it for the routing branch -- the knob BASELINE.json's
        # iteration sweep turns -- and is ignored when absent, so the reference's own params files behave as before
        p_iter = getattr(params, 'n_iter', None) if n_nodes != -1 else None
        self.n_iter = int(p_iter) if p_iter is not None else n_iter
        self.n_nodes = n_nodes
        self.n_caps = n_caps
        if n_nodes != -1:   # caps -> caps: the routing branch (models.py:56-58)
            self.route_weights = nn.Parameter(0.1 * torch.randn(1, n_nodes, n_caps, in_C, out_C))
        else:               # conv -> caps: primary capsules (models.py:59-62), same modules as the reference
            self.capsules = nn.ModuleList(
                [nn.Conv2d(in_C, out_C, kernel, stride=stride) for _ in range(n_caps)])

    def squash(self, v):
        if v.is_cuda and v.dtype == torch.float32:
            return _SquashFn.apply(v)
        sq = (v ** 2).sum(dim=-1, keepdim=True)
        return (sq / (1 + sq)) * v / torch.sqrt(sq)

    def forward(self, x):
        if self.n_nodes != -1:
            v = dynamic_routing(x, self.route_weights, self.n_iter)       # [B,C,D]
            B, C, D = v.shape
            return v.view(B, 1, C, 1, D)                                   # reference output shape
        if x.is_cuda and x.dtype == torch.float32 and len(self.capsules) <= 16:
            return primary_capsules(x, self.capsules)
        outs = [cap(x).view(x.size(0), -1, 1) for cap in self.capsules]     # reference models.py:81-82
        return self.squash(torch.cat(outs, dim=-1))

    def forward_margin_loss(self, x, y):
        """Fused variant: ([B,1,C,1,D] output, margin loss); see routing_margin_loss."""
        v, loss = routing_margin_loss(x, self.route_weights, y, self.n_iter)
        B, C, D = v.shape
        return v.view(B, 1, C, 1, D), loss


class HostPipe:
    """Double-buffered end-to-end call (caps_host_pipe_*): `submit` starts the host->device copy of a batch on an
    internal stream, `step` runs forward + margin loss + fused backward on the batch submitted first and returns the
    loss on the host.  Submitting batch n+1 before stepping batch n hides the copy under the kernels."""

    def __init__(self, B, N, C, K, D, n_iter, device='cuda'):
        import ctypes
        L = _cabi.lib()
        self.dims = (B, N, C, K, D, n_iter)
        nbytes = L.caps_host_pipe_scratch_bytes(B, N, C, K, D, n_iter)
        if nbytes == 0:
            raise RuntimeError('unsupported dims %s' % (self.dims,))
        self.device = torch.device(device)
        self.scratch = torch.empty((nbytes,), device=self.device, dtype=torch.uint8)
        self.loss_host = torch.empty((1,), dtype=torch.float32).pin_memory()
        self._pipe = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _cabi.check(L.caps_host_pipe_create(ctypes.byref(self._pipe), _ptr(self.scratch), nbytes, B, N, C, K, D, n_iter),
                        'caps_host_pipe_create')
        self.h2d_bytes = B * N * K * 4 + B * 8
        self.d2h_bytes = 4

    def submit(self, u_host, y_host):
        with torch.cuda.device(self.device):
            _cabi.check(_cabi.lib().caps_host_pipe_submit(self._pipe, _ptr(u_host), _ptr(y_host)), 'caps_host_pipe_submit')

    def step(self, W_dev, dW_dev, v_host=None, dw_ready_event=None):
        ev = None if dw_ready_event is None else dw_ready_event.cuda_event
        with torch.cuda.device(self.device):
            _cabi.check(_cabi.lib().caps_host_pipe_step(self._pipe, _ptr(W_dev), _ptr(dW_dev), _ptr(self.loss_host), _ptr(v_host),
                                                        _stream(), ev), 'caps_host_pipe_step')
        return self.loss_host

    def close(self):
        if self._pipe:
            _cabi.lib().caps_host_pipe_destroy(self._pipe)
            self._pipe = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class HostStep:
    """End-to-end call with HOST buffers through caps_route_step_host: u, y come from (pinned)
    host memory every step, the loss (and optionally v / du) go back to the host; W and dW stay
    on the device like the reference's parameters do."""

    def __init__(self, B, N, C, K, D, n_iter, device='cuda'):
        L = _cabi.lib()
        self.dims = (B, N, C, K, D, n_iter)
        nbytes = L.caps_route_step_host_scratch_bytes(B, N, C, K, D, n_iter)
        if nbytes == 0:
            raise RuntimeError('unsupported dims %s' % (self.dims,))
        self.device = torch.device(device)
        self.scratch = torch.empty((nbytes,), device=self.device, dtype=torch.uint8)
        self.loss_host = torch.empty((1,), dtype=torch.float32).pin_memory()
        self.h2d_bytes = B * N * K * 4 + B * 8
        self.d2h_bytes = 4 if B < 2048 else 12          # one loss term per micro-batch

    def __call__(self, u_host, y_host, W_dev, dW_dev, v_host=None, du_host=None):
        B, N, C, K, D, R = self.dims
        with torch.cuda.device(self.device):
            _cabi.check(_cabi.lib().caps_route_step_host(
                _ptr(u_host), _ptr(y_host), _ptr(W_dev), _ptr(self.loss_host), _ptr(v_host), _ptr(du_host),
                _ptr(dW_dev), _ptr(self.scratch), self.scratch.numel(), B, N, C, K, D, R, _stream()),
                'caps_route_step_host')
        return self.loss_host


class GraphedStep:
    """One routing train step -- forward, margin loss, fused backward: the 17 launches of caps_route_forward /
    caps_margin_loss / caps_route_backward -- captured ONCE in a CUDA graph over static device buffers and replayed
    with a single launch.  For the reference's own operating point (experiments/capsule/params.json: batch 64) the
    step is launch-bound: the kernels are ~20 us each and the graph removes the gaps between them.

        g = GraphedStep(B, N, C, K, D, n_iter, W)       # W [N,C,K,D] device tensor: read in place at every replay
        g.u.copy_(u); g.y.copy_(y); g.replay()           # -> g.v, g.loss, g.du, g.dW
    """

    def __init__(self, B, N, C, K, D, n_iter, W, device=None):
        L = _cabi.lib()
        dev = W.device if device is None else torch.device(device)
        self.dims = (B, N, C, K, D, n_iter)
        self.W = W
        self.u = torch.zeros(B, N, K, device=dev)
        self.y = torch.zeros(B, dtype=torch.int64, device=dev)
        self.v = torch.empty(B, C, D, device=dev)
        self.du = torch.empty(B, N, K, device=dev)
        self.dW = torch.empty(N, C, K, D, device=dev)
        self.loss = torch.empty((), device=dev)
        self._lscr = torch.empty(_cabi.MARGIN_SCRATCH_FLOATS, device=dev)
        nbytes = L.caps_route_workspace_bytes(B, N, C, K, D, n_iter, 1)
        if nbytes == 0:
            raise RuntimeError('unsupported dims %s' % (self.dims,))
        self._ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        self._nbytes = nbytes
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.device(dev):
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                self.u.normal_()                      # warm-up on real-looking data (also sets the kernel attributes,
                self._enqueue()                       # which must not happen inside a capture)
                self._enqueue()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            with torch.cuda.graph(self.graph):
                self._enqueue()

    def _enqueue(self):
        L = _cabi.lib()
        B, N, C, K, D, R = self.dims
        st = _stream()
        _cabi.check(L.caps_route_forward(_ptr(self.u), _ptr(self.W), _ptr(self.v), None, _ptr(self._ws), self._nbytes,
                                         B, N, C, K, D, R, 1, st), 'caps_route_forward')
        _cabi.check(L.caps_margin_loss(_ptr(self.v), _ptr(self.y), 1.0 / B, _ptr(self.loss), None, _ptr(self._lscr), B, C, D, st),
                    'caps_margin_loss')
        _cabi.check(L.caps_route_backward(_ptr(self.u), _ptr(self.W), None, _ptr(self.y), 1.0 / B, None, _ptr(self.du), _ptr(self.dW),
                                          _ptr(self._ws), self._nbytes, B, N, C, K, D, R, st), 'caps_route_backward')

    def replay(self):
        self.graph.replay()
        return self.loss
