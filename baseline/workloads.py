"""BASELINE.json's configs as runnable steps, for three arms:

  reference/cpu   the unmodified reference (baseline/_ref) on the host cores
  reference/cuda  the same unmodified reference ops run eagerly on the B200 (informational baseline: BASELINE.md section 3)
  dropin/cuda     the reference's own model / loss / train-step code with `models.CapsuleLayer` replaced by the
                  B200-native layer before the model is built, optionally with the one-kernel regroup /
                  loss tail of section 8(f) for DarkCapsuleNet

A step is what reference main.py:55-77 does per batch: H2D of the numpy batch, forward, loss, `.cpu().numpy()` of the
prediction, zero_grad / backward / Adam step, `loss.item()`.  Synthetic data has the shapes and value ranges of
SURVEY 8(d).  Measurement infrastructure only (bench.py, tests)."""
import os
import time

import numpy as np
import torch

from . import refload

N_CLASSES = 43


# ------------------------------------------------------------------------------------------------------------------
# synthetic data
# ------------------------------------------------------------------------------------------------------------------
def synth_cfg1(B, seed=0):
    """GTSRB-shaped: x NHWC float32 in [-1, 1) (utils.center_rgb range), y int64 labels."""
    rng = np.random.default_rng(seed)
    x = (rng.random((B, 32, 32, 3), dtype=np.float32) * 2 - 1).astype(np.float32)
    y = rng.integers(0, N_CLASSES, size=(B,), dtype=np.int64)
    return x, y


def synth_cfg3(B, seed=0, grid=7):
    """GTSDB-shaped frames at the size the model accepts (224, config.py:41): about one object cell per image;
    label = (objectness, x, y, w, h, one-hot class) per cell."""
    rng = np.random.default_rng(seed)
    x = (rng.random((B, 224, 224, 3), dtype=np.float32) * 2 - 1).astype(np.float32)
    y = np.zeros((B, grid, grid, 5 + N_CLASSES), dtype=np.float32)
    y[..., 1:5] = rng.uniform(0.05, 0.95, size=(B, grid, grid, 4))
    gy, gx = rng.integers(0, grid, size=B), rng.integers(0, grid, size=B)
    cls = rng.integers(0, N_CLASSES, size=B)
    y[np.arange(B), gy, gx, 0] = 1.0
    y[np.arange(B), gy, gx, 5 + cls] = 1.0
    return x, y


def synth_routing(B, N, C=N_CLASSES, K=8, D=16, seed=0):
    """cfg2: u = squash(N(0,1)) [B,N,K], W = 0.1 N(0,1) [N,C,K,D], y labels."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, N, K, generator=g)
    sq = (x ** 2).sum(-1, keepdim=True)
    u = (sq / (1 + sq)) * x / sq.sqrt()
    W = 0.1 * torch.randn(N, C, K, D, generator=torch.Generator().manual_seed(99))
    y = torch.randint(0, C, (B,), generator=g)
    return u.contiguous(), W.contiguous(), y


# ------------------------------------------------------------------------------------------------------------------
# model builders
# ------------------------------------------------------------------------------------------------------------------
def _build(ref, cls_name, params, device, dropin_layer):
    """Builds a reference model; with dropin_layer the class name `CapsuleLayer` is rebound while __init__ runs
    (the reference resolves it at construction time) and restored afterwards."""
    orig = ref.models.CapsuleLayer
    try:
        if dropin_layer is not None:
            ref.models.CapsuleLayer = dropin_layer
        torch.manual_seed(0)
        model = getattr(ref.models, cls_name)(params)
    finally:
        ref.models.CapsuleLayer = orig
    return model.to(device)


def _to_nchw(x_np, device):
    """main.py:57: `torch.from_numpy(x).float().permute(0, 3, 1, 2).to(device)`.  Under torch >= 1.5 the permuted view is
    a channels-last tensor, convolutions keep that memory format, and the reference's `.view` in models.py:81 then
    raises (torch 0.4, which the reference pins, always produced NCHW-contiguous outputs).  `.contiguous()` restores
    the layout the reference was written against; it is the only deviation from main.py's step, on every arm."""
    return torch.from_numpy(x_np).float().permute(0, 3, 1, 2).contiguous().to(device=device)


def make_cfg1_step(device, dropin=False):
    """BASELINE.json configs[0]: CapsNet train step, recon on, Adam (reference main.py:55-77, models.py:113-124,
    loss_fns.py:11-23).  Returns (step(x_np, y_np) -> loss float, model)."""
    ref = refload.load()
    if ref is None:
        raise RuntimeError('reference not installed (baseline/_ref)')
    layer = None
    if dropin:
        import cs231_capsule_yolo_traffic_sign_detection_b200 as pkg
        layer = pkg.CapsuleLayer
    params = refload.make_params(ref, 'capsule', str(device), recon=True)
    model = _build(ref, 'CapsuleNet', params, device, layer)
    model.train()
    opt = torch.optim.Adam(filter(lambda p: p.requires_grad, model.parameters()), lr=1e-3)

    def step(x_np, y_np):
        x = _to_nchw(x_np, device)                                                     # main.py:57-59
        y = torch.from_numpy(y_np).to(device=device)
        y_hat, recon = model(x, y, True)                                               # main.py:62
        loss = ref.loss_fns.capsule_loss(y_hat, y, params, x, recon)                   # main.py:63
        _ = y_hat.data.cpu().numpy()                                                   # main.py:68 (a sync)
        opt.zero_grad()
        loss.backward()
        opt.step()
        return loss.item()                                                             # main.py:74 (second sync)
    return step, model


def make_cfg1_epoch(device, batch, graph=False):
    """configs[0] through the sync-free side runner (package runner.py): the same model, loss and
    optimizer as make_cfg1_step(dropin=True), but a whole epoch of batches per call -- staged copies one batch ahead,
    no per-step host synchronisation.  Returns epoch(x_np, y_np) -> avg_loss."""
    ref = refload.load()
    if ref is None:
        raise RuntimeError('reference not installed (baseline/_ref)')
    import cs231_capsule_yolo_traffic_sign_detection_b200 as pkg
    params = refload.make_params(ref, 'capsule', str(device), recon=True)
    params.batch_size = batch
    model = _build(ref, 'CapsuleNet', params, device, pkg.CapsuleLayer)
    opt = torch.optim.Adam(filter(lambda p: p.requires_grad, model.parameters()), lr=1e-3, capturable=graph)

    def epoch(x_np, y_np):
        return pkg.runner.train(x_np, y_np, model, opt, ref.loss_fns.capsule_loss, None, params, no_metric=True, shuffle=False, graph=graph)[0]
    return epoch


def make_cfg3_step(device, dropin=False, fused_tail=False, grid=7, world=1):
    """BASELINE.json configs[2]: DarkCapsuleNet train step with `--recon --no_metric` semantics (reference
    main.py:55-77, models.py:389-400, loss_fns.py:187-204).  fused_tail: the cell regroup and the loss run as
    the one-kernel replacements of SURVEY 8(f) rows 2 and 3 instead of the reference's op chains."""
    ref = refload.load()
    if ref is None:
        raise RuntimeError('reference not installed (baseline/_ref)')
    pkg = None
    if dropin:
        import cs231_capsule_yolo_traffic_sign_detection_b200 as pkg
    params = refload.make_params(ref, 'darkcapsule', str(device), recon=False)
    params.n_grid = grid
    model = _build(ref, 'DarkCapsuleNet', params, device, pkg.CapsuleLayer if dropin else None)
    model.train()
    opt = torch.optim.Adam(filter(lambda p: p.requires_grad, model.parameters()), lr=1e-3)
    bucket = None
    if world > 1:
        from cs231_capsule_yolo_traffic_sign_detection_b200.parallel import GradBucket
        # the decoder is built but never used by this forward (models.py:389-400): its parameters get no gradient
        used = [p for n, p in model.named_parameters() if not n.startswith('decoder.')]
        bucket = GradBucket(used)

    def step(x_np, y_np):
        x = _to_nchw(x_np, device)
        y = torch.from_numpy(y_np).to(device=device)
        if fused_tail:
            out = model.traffic_sign_capsules(pkg.dark_regroup(model.conv(x), grid))      # [g*g*B,1,1,1,5]
            loss = pkg.dark_capsule_loss(out, y)
            y_hat = out.view(grid, grid, x.size(0), 5).permute(2, 0, 1, 3)                # models.py:399
        else:
            y_hat = model(x)
            loss = ref.loss_fns.darkcapsule_loss(y_hat, y, params)
        _ = y_hat.data.cpu().numpy()
        if bucket is not None:
            bucket.zero()
        else:
            opt.zero_grad()
        loss.backward()
        if bucket is not None:
            bucket.allreduce(average=True)
        opt.step()
        return loss.item()
    return step, model


def make_routing_reference_step(device, N, C=N_CLASSES, K=8, D=16, R=3):
    """cfg2 on the unmodified reference: CapsuleLayer routing branch + capsule_loss (recon off) + backward."""
    ref = refload.load()
    if ref is None:
        raise RuntimeError('reference not installed (baseline/_ref)')
    params = refload.make_params(ref, 'capsule', str(device), recon=False)
    params.n_classes = C
    torch.manual_seed(0)
    layer = ref.models.CapsuleLayer(params, n_caps=C, n_nodes=N, in_C=K, out_C=D, n_iter=R).to(device)

    def step(u, W, y):
        """u [B,N,K], W [N,C,K,D], y [B] already on `device`; returns (loss, dW) like one training step would."""
        with torch.no_grad():
            layer.route_weights.copy_(W[None])
        layer.route_weights.grad = None
        uu = u.detach().requires_grad_(True)
        out = layer(uu).squeeze()                                  # models.py:116
        scores = (out ** 2).sum(dim=-1) ** 0.5                     # models.py:117
        loss = ref.loss_fns.capsule_loss(scores, y, params)        # loss_fns.py:11-23
        loss.backward()
        return loss, layer.route_weights.grad, uu.grad
    return step


# ------------------------------------------------------------------------------------------------------------------
# timing
# ------------------------------------------------------------------------------------------------------------------
def time_steps(fn, steps, warmup, cuda):
    """Seconds per step: wall clock bracketed by device syncs (every step already ends in a host sync, loss.item())."""
    for _ in range(warmup):
        fn()
    if cuda:
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    if cuda:
        torch.cuda.synchronize()
    return (time.perf_counter() - t0) / max(steps, 1)


def cpu_threads():
    n = os.cpu_count() or 1
    torch.set_num_threads(n)
    return torch.get_num_threads()
