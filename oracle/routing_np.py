"""CPU oracle (numpy) for the capsule dynamic-routing hot path.  TEST INFRASTRUCTURE ONLY.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this package.  The product path (`cs231_capsule_yolo_traffic_sign_detection_b200`)
never does, and fails loudly when its CUDA library is missing.

This file is a closed-form restatement (explicit forward AND explicit backward, no autograd)
of the caps->caps branch of the reference `CapsuleLayer`:

    reference models.py:64-67   squash
    reference models.py:70-79   prediction vectors + routing loop
    reference models.py:81-82   primary-capsule tail (views + cat + squash), the step before the routing layer
    reference models.py:393-399 DarkCapsuleNet cell regroup, the step before the routing layer in that model
    reference loss_fns.py:187-204, utils.py:69-85  darkcapsule_loss + polar_transform, the step after it
    reference models.py:116-117 class scores (norm of the class capsules)
    reference loss_fns.py:11-17,23  margin loss (recon term excluded: it is not on the path)

Parity pin: the reference ships no golden vectors / tests, so this oracle is
pinned against outputs of the reference itself, run in the build container by
`tests/golden/make_golden.py` (imports /root/reference/models.py unmodified) and committed as
`tests/golden/*.npz`; `tests/test_oracle.py` checks this file against those fixtures.

Shapes:  u [B,N,K]   W [N,C,K,D]   v [B,C,D]   c [B,N,C]   y [B] int
"""
import numpy as np


def squash(s):
    """reference models.py:64-67 -- v = (|s|^2/(1+|s|^2)) * s / sqrt(|s|^2); no epsilon."""
    n2 = (s * s).sum(-1, keepdims=True)
    return (n2 / (1.0 + n2)) * s / np.sqrt(n2)


def squash_bwd(s, dv):
    """d/ds of squash: with f(n2) = sqrt(n2)/(1+n2), v = f*s,
    ds = f*dv + s * (s.dv) * (1-n2) / (sqrt(n2) (1+n2)^2)."""
    n2 = (s * s).sum(-1, keepdims=True)
    n = np.sqrt(n2)
    sdv = (s * dv).sum(-1, keepdims=True)
    return dv * n / (1.0 + n2) + s * sdv * (1.0 - n2) / (n * (1.0 + n2) ** 2)


def primary_tail(conv, n_caps):
    """reference models.py:81-82 -- `[cap(x).view(B,-1,1) for cap in capsules]`, cat(dim=-1), squash.
    conv [B, n_caps*Cc, H, W] holds the n_caps conv outputs capsule-major (channel = k*Cc + c);
    returns u [B, Cc*H*W, n_caps]."""
    B, KC, H, W = conv.shape
    pre = conv.reshape(B, n_caps, (KC // n_caps) * H * W).transpose(0, 2, 1)
    return squash(pre)


def primary_tail_bwd(conv, du, n_caps):
    """gradient of primary_tail w.r.t. conv, same layout as conv."""
    B, KC, H, W = conv.shape
    pre = conv.reshape(B, n_caps, (KC // n_caps) * H * W).transpose(0, 2, 1)
    return squash_bwd(pre, du).transpose(0, 2, 1).reshape(conv.shape)


def dark_regroup(x, n_grid):
    """reference models.py:393-399 -- view(B,Cch,4,4g^2), chunk(g^2, dim 3), per chunk
    permute(0,2,3,1).contiguous().view(B,-1,8), stacked cell-major: [g^2*B, 2*Cch, 8]."""
    B, Cch = x.shape[0], x.shape[1]
    G = n_grid * n_grid
    xv = x.reshape(B, Cch, 4, G, 4)                        # [b, ch, a, q, t]
    return np.ascontiguousarray(xv.transpose(3, 0, 2, 4, 1)).reshape(G * B, 2 * Cch, 8)   # [q, b, a, t, ch]


def dark_regroup_bwd(du, B, Cch, n_grid):
    """inverse map: gradient w.r.t. x [B, Cch, 16 g^2] from du [g^2*B, 2*Cch, 8]."""
    G = n_grid * n_grid
    dv = du.reshape(G, B, 4, 4, Cch)                       # [q, b, a, t, ch]
    return np.ascontiguousarray(dv.transpose(1, 4, 2, 0, 3)).reshape(B, Cch, 16 * G)      # [b, ch, a, q, t]


def polar_transform(x):
    """reference utils.py:69-85 -- (r, x, y, w, h) -> r and the 5-d direction vector."""
    r, xx, yy, w, h = [x[..., k] for k in range(5)]
    f1, f2, f3, f4 = xx * np.pi, yy * np.pi, h * np.pi, w * np.pi * 2
    s1, c1, s2, c2, s3, c3, s4, c4 = np.sin(f1), np.cos(f1), np.sin(f2), np.cos(f2), np.sin(f3), np.cos(f3), np.sin(f4), np.cos(f4)
    return r, np.stack([s1, s1 * c2, s1 * s2 * c3, s1 * s2 * s3 * c4, s1 * s2 * s3 * s4], axis=-1)


def dark_loss(caps, y):
    """reference loss_fns.py:187-204 (recon off): caps [B,g,g,5], y [B,g,g,>=5] -> (loss, d loss / d caps)."""
    y_r, y_phi = polar_transform(y[..., :5])
    m = np.sqrt((caps * caps).sum(-1))
    left, right = np.maximum(0.9 - m, 0.0), np.maximum(m - 0.1, 0.0)
    B = y.shape[0]
    loss = ((y_r * left ** 2 + 0.5 * (1 - y_r) * right ** 2).sum() - (caps * y_phi).sum()) / B
    dm = (-2.0 * y_r * left + (1 - y_r) * right) / m
    return loss, (dm[..., None] * caps - y_phi) / B


def softmax_c(b):
    """reference models.py:75 -- softmax over the class-capsule axis (max-subtracted)."""
    e = np.exp(b - b.max(-1, keepdims=True))
    return e / e.sum(-1, keepdims=True)


def prediction_vectors(u, W):
    """reference models.py:71 -- u_hat[b,i,j,:] = u[b,i,:] @ W[i,j,:,:]."""
    return np.einsum('bik,ijkd->bijd', u, W)


def routing_forward(u, W, n_iter=3, return_state=False):
    """reference models.py:70-79.  Returns v [B,C,D] (and, optionally, the per-iteration state).

    Uses the identity logits^r = u_hat . (v^0 + ... + v^{r-1}) so logits are [B,N,C]
    (the reference keeps them D-redundant as [B,N,C,1,D])."""
    uh = prediction_vectors(u, W)                      # [B,N,C,D]
    B, N, C, D = uh.shape
    b = np.zeros((B, N, C), dtype=uh.dtype)            # models.py:72
    s_list, v_list, c_list = [], [], []
    for r in range(n_iter):
        c = softmax_c(b)                               # models.py:75
        s = np.einsum('bij,bijd->bjd', c, uh)          # models.py:76 (inner)
        v = squash(s)                                  # models.py:76
        s_list.append(s); v_list.append(v); c_list.append(c)
        if r != n_iter - 1:
            b = b + np.einsum('bijd,bjd->bij', uh, v)  # models.py:78-79
    if return_state:
        return v, dict(uh=uh, s=s_list, v=v_list, c=c_list)
    return v


def routing_backward(u, W, grad_v, n_iter=3, state=None):
    """Explicit reverse pass through all routing iterations (the reference never detaches).

    Returns (du [B,N,K], dW [N,C,K,D]).  Follows SURVEY.md section 7.1:
      beta^r  = total gradient w.r.t. logits b^r (carried through the identity path),
      G_bij   = sum_r c^r_ij ds^r_j + sum_{r>=1} beta^r_ij v^{r-1}_j   (= d loss / d u_hat_bij)."""
    if state is None:
        _, state = routing_forward(u, W, n_iter, return_state=True)
    uh, s_l, v_l, c_l = state['uh'], state['s'], state['v'], state['c']
    B, N, C, D = uh.shape
    G = np.zeros_like(uh)
    beta = np.zeros((B, N, C), dtype=uh.dtype)
    dv = grad_v
    for r in range(n_iter - 1, -1, -1):
        ds = squash_bwd(s_l[r], dv)                              # [B,C,D]
        G += c_l[r][..., None] * ds[:, None]
        if r > 0:
            dc = np.einsum('bijd,bjd->bij', uh, ds)
            t = (c_l[r] * dc).sum(-1, keepdims=True)
            beta = beta + c_l[r] * (dc - t)                      # softmax backward + carry
            G += beta[..., None] * v_l[r - 1][:, None]
            dv = np.einsum('bij,bijd->bjd', beta, uh)            # grad w.r.t. v^{r-1}
    dW = np.einsum('bik,bijd->ijkd', u, G)
    du = np.einsum('ijkd,bijd->bik', W, G)
    return du, dW


def class_scores(v):
    """reference models.py:117 -- scores = sqrt(sum_d v^2)."""
    return np.sqrt((v * v).sum(-1))


def margin_loss(v, y, inv_batch=None):
    """reference loss_fns.py:12-17,23 on scores = |v| (recon term off).  Returns a scalar."""
    B, C, _ = v.shape
    if inv_batch is None:
        inv_batch = 1.0 / B
    m = class_scores(v)
    T = np.zeros((B, C), dtype=v.dtype)
    T[np.arange(B), y] = 1.0
    left = np.maximum(0.9 - m, 0.0) ** 2
    right = np.maximum(m - 0.1, 0.0) ** 2
    return (T * left + 0.5 * (1.0 - T) * right).sum() * inv_batch


def margin_loss_grad(v, y, inv_batch=None):
    """d margin_loss / d v  [B,C,D]."""
    B, C, _ = v.shape
    if inv_batch is None:
        inv_batch = 1.0 / B
    m = class_scores(v)
    T = np.zeros((B, C), dtype=v.dtype)
    T[np.arange(B), y] = 1.0
    dm = (-2.0 * T * np.maximum(0.9 - m, 0.0) + (1.0 - T) * np.maximum(m - 0.1, 0.0)) * inv_batch
    return dm[..., None] * v / m[..., None]


def routing_step(u, W, y, n_iter=3, grad_v_extra=None):
    """fwd + margin loss + bwd: what one 'step' of the hot path computes.
    Returns dict(v, c, loss, du, dW)."""
    v, st = routing_forward(u, W, n_iter, return_state=True)
    loss = margin_loss(v, y)
    g = margin_loss_grad(v, y)
    if grad_v_extra is not None:
        g = g + grad_v_extra
    du, dW = routing_backward(u, W, g, n_iter, st)
    return dict(v=v, c=st['c'][-1], loss=loss, du=du, dW=dW)


def make_inputs(B, N, C, K, D, seed=0, dtype=np.float32):
    """Seeded synthetic inputs shaped like the reference's:
    u = squash(N(0,1)) per primary capsule (what models.py:82 emits), W = 0.1*N(0,1)
    (models.py:57-58), y ~ randint(C)."""
    rng = np.random.default_rng(seed)
    u = squash(rng.standard_normal((B, N, K))).astype(dtype)
    W = (0.1 * rng.standard_normal((N, C, K, D))).astype(dtype)
    y = rng.integers(0, C, size=(B,)).astype(np.int64)
    return u, W, y
