"""CPU oracle (torch, op-for-op) for the capsule dynamic-routing hot path.  TEST INFRASTRUCTURE ONLY.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs
may import this.  It is never on the product path.

The reference's hot path *is* a chain of stock ATen ops; this file restates
that chain with the same ops in the same order so that, on the same host, it costs what the
reference costs and rounds the way the reference rounds:

    batched matmul of [B,N,1,1,K] by [1,N,C,K,D]          reference models.py:71
    zero logits the size of the priors (D-redundant)      reference models.py:72
    softmax over dim 2 / weighted sum over dim 1 / squash reference models.py:75-76, 64-67
    agreement = sum over D, broadcast-added to logits     reference models.py:77-79
    scores = L2 norm over D                               reference models.py:117
    margin loss                                           reference loss_fns.py:12-17,23

The backward is torch autograd, exactly as in the reference (main.py:71).
Parity pin: `tests/test_oracle.py` checks this port against fixtures written by the unmodified
reference (`tests/golden/make_golden.py`); on the same torch build they agree bit for bit.
"""
import torch
import torch.nn.functional as F


def squash_t(x):
    # reference models.py:64-67
    sq = (x ** 2).sum(dim=-1, keepdim=True)
    return (sq / (1 + sq)) * x / torch.sqrt(sq)


def routing_forward_t(u, W5, n_iter=3, want_c=False):
    """u [B,N,K]; W5 [1,N,C,K,D] (the reference's parameter shape).  Returns [B,1,C,1,D]
    (and the last coupling coefficients [B,N,C] when want_c)."""
    uhat = (u[:, :, None, None, :] @ W5).squeeze(4)            # [B,N,C,1,D]
    blog = torch.zeros(*uhat.size(), dtype=uhat.dtype)
    out, coup = None, None
    for it in range(n_iter):
        coup = F.softmax(blog, dim=2)
        out = squash_t((coup * uhat).sum(dim=1, keepdim=True))
        if it != n_iter - 1:
            blog = blog + (uhat * out).sum(dim=-1, keepdim=True)
    if want_c:
        return out, coup[:, :, :, 0, 0]
    return out


def margin_loss_t(scores, y, n_classes):
    # reference loss_fns.py:12-17,23 with params.recon False
    pos = F.relu(0.9 - scores) ** 2
    neg = F.relu(scores - 0.1) ** 2
    onehot = torch.eye(n_classes, dtype=scores.dtype).index_select(dim=0, index=y)
    return (onehot * pos + 0.5 * (1. - onehot) * neg).sum() / y.size(0)


def routing_step_t(u, W5, y, n_iter=3, want_c=False):
    """One hot-path step on CPU: fwd + scores + margin loss + autograd bwd.
    Returns dict(v [B,C,D], c, loss, du, dW [N,C,K,D])."""
    u = u.detach().clone().requires_grad_(True)
    W5 = W5.detach().clone().requires_grad_(True)
    res = routing_forward_t(u, W5, n_iter, want_c)
    out, c = res if want_c else (res, None)
    C, D = out.shape[2], out.shape[4]
    v = out.reshape(out.shape[0], C, D)                        # == .squeeze() for B>1, C>1
    scores = (v ** 2).sum(dim=-1) ** 0.5                       # reference models.py:117
    loss = margin_loss_t(scores, y, C)
    loss.backward()
    return dict(v=v.detach(), c=None if c is None else c.detach(), loss=loss.detach(),
                du=u.grad, dW=W5.grad[0])
