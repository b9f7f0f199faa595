"""Generate the golden fixtures in this directory by running the UNMODIFIED reference.

Run in the build container only (it needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

The reference's `models.py` / `loss_fns.py` are imported as they are; the only shim is a stub for
`matplotlib.colors` (not installed here; used only by `utils.augmentation`, which the training
loop has commented out, reference main.py:56).  Inputs come from
`oracle.routing_np.make_inputs(seed)` so that tests can regenerate them bit-identically on any
box; each fixture stores an input checksum to detect RNG drift.

Per case the fixture holds what the reference layer + its own loss + autograd produced in fp32
(and in fp64, as the tie-breaker for tolerance budgeting):
  v [B,C,D], c_last [B,N,C] (last-iteration softmax, captured by wrapping F.softmax),
  loss (capsule_loss with recon off), du [B,N,K], dW [N,C,K,D]  (or a strided probe of dW for
  the full-size case, to keep the fixture small).
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = '/root/reference'

# name: (B, N, C, K, D, R, seed, full_dW)
CASES = {
    'caps_c43_d16_r3':   (4, 12, 43, 8, 16, 3, 11, True),    # CapsuleNet head, shrunk in N
    'caps_c10_d16_r3':   (7, 40, 10, 8, 16, 3, 12, True),
    'caps_c43_d16_r1':   (3, 8, 43, 8, 16, 1, 13, True),    # single iteration (no agreement step)
    'caps_c43_d16_r5':   (3, 8, 43, 8, 16, 5, 14, True),
    'dark_c1_d5_r3':     (9, 512, 1, 8, 5, 3, 15, True),     # DarkCapsuleNet head (models.py:368-370)
    'dark3_c43_d21_r3':  (2, 6, 43, 8, 21, 3, 16, True),    # DarkCapsuleNet3 head shape (models.py:431-433)
    'dark2_c49_d48_r3':  (2, 4, 49, 8, 48, 3, 17, True),     # DarkCapsuleNet2 head shape (models.py:327-329)
    'sweep_c43_d32_r2':  (5, 6, 43, 8, 32, 2, 18, True),
    'caps_b1':           (1, 8, 43, 8, 16, 3, 19, True),    # batch of one
    'caps_full_n1296':   (2, 1296, 43, 8, 16, 3, 20, False), # the real CapsuleNet routing shape
    'caps_full_n1152':   (3, 1152, 43, 8, 16, 3, 21, False), # BASELINE.json config 2 shape
}
PROBE_STRIDE = 997


def import_reference():
    mpl = types.ModuleType('matplotlib')
    col = types.ModuleType('matplotlib.colors')
    col.rgb_to_hsv = lambda a: a
    col.hsv_to_rgb = lambda a: a
    mpl.colors = col
    sys.modules.setdefault('matplotlib', mpl)
    sys.modules.setdefault('matplotlib.colors', col)
    sys.path.insert(0, REF)
    import models as ref_models      # noqa: E402
    import loss_fns as ref_loss      # noqa: E402
    return ref_models, ref_loss


class P:  # stand-in for utils.Params (an attribute bag, reference utils.py:14-31)
    device = 'cpu'
    recon = False
    recon_coef = 5e-4


def run_reference(ref_models, ref_loss, u, W, y, R, dtype):
    B, N, K = u.shape
    _, C, _, D = W.shape
    params = P()
    params.n_classes = C
    layer = ref_models.CapsuleLayer(params, n_caps=C, n_nodes=N, in_C=K, out_C=D, n_iter=R)
    with torch.no_grad():
        layer.route_weights.copy_(torch.from_numpy(W)[None])
    layer = layer.to(dtype)
    ut = torch.from_numpy(u).to(dtype).requires_grad_(True)

    captured = []
    real_softmax = ref_models.F.softmax

    def spy(x, dim):
        out = real_softmax(x, dim=dim)
        captured.append(out)
        return out
    ref_models.F.softmax = spy
    try:
        out = layer(ut)                                    # [B,1,C,1,D]
    finally:
        ref_models.F.softmax = real_softmax
    v = out.reshape(B, C, D)
    scores = (v ** 2).sum(dim=-1) ** 0.5                   # reference models.py:117
    loss = ref_loss.capsule_loss(scores, torch.from_numpy(y), params)
    loss.backward()
    c_last = captured[-1][:, :, :, 0, 0]
    return dict(v=v.detach().numpy(), c=c_last.detach().numpy(), loss=float(loss.detach()),
                du=ut.grad.numpy(), dW=layer.route_weights.grad[0].numpy())


def make_primary(ref_models):
    """Primary-capsule branch (reference models.py:59-62, 81-82), the step before the routing layer:
    K convolutions -> view -> cat(dim=-1) -> squash.  Stores inputs, the K conv outputs stacked
    capsule-major, the layer output, and autograd's gradients for a seeded upstream gradient."""
    B, Cin, H, K, Cc, kern, stride = 3, 6, 14, 8, 5, 4, 2
    g = torch.Generator().manual_seed(31)
    layer = ref_models.CapsuleLayer(P(), n_caps=K, n_nodes=-1, in_C=Cin, out_C=Cc, kernel=kern, stride=stride)
    with torch.no_grad():
        for m in layer.capsules:
            m.weight.copy_(0.2 * torch.randn(m.weight.shape, generator=g))
            m.bias.copy_(0.1 * torch.randn(m.bias.shape, generator=g))
    x = torch.randn(B, Cin, H, H, generator=g).requires_grad_(True)
    outs = []
    def keep(mod, inp, out):
        out.retain_grad()
        outs.append(out)
    hooks = [m.register_forward_hook(keep) for m in layer.capsules]
    u = layer(x)                                              # [B, Cc*H'*W', K]
    for h in hooks:
        h.remove()
    du = torch.randn(u.shape, generator=g)
    u.backward(du)
    conv = torch.cat(outs, dim=1)                             # [B, K*Cc, H', W'], channel = k*Cc + c
    dconv = torch.cat([o.grad for o in outs], dim=1)
    blob = dict(dims=np.array([B, Cin, H, K, Cc, kern, stride], dtype=np.int64),
                x=x.detach().numpy(), weight=np.stack([m.weight.detach().numpy() for m in layer.capsules]),
                bias=np.stack([m.bias.detach().numpy() for m in layer.capsules]),
                conv=conv.detach().numpy(), u=u.detach().numpy(), du=du.numpy(), dconv=dconv.numpy(),
                dx=x.grad.numpy(), dweight=np.stack([m.weight.grad.numpy() for m in layer.capsules]),
                dbias=np.stack([m.bias.grad.numpy() for m in layer.capsules]))
    path = os.path.join(HERE, 'primary_caps.npz')
    np.savez_compressed(path, **blob)
    print('%-20s conv %s -> u %s -> %s (%d KB)' % ('primary_caps', tuple(conv.shape), tuple(u.shape),
                                                os.path.basename(path), os.path.getsize(path) // 1024))


def dark_pattern(shape, mul, mod):
    """integer-valued fp32 test pattern (exact through any permutation, compresses well)"""
    n = int(np.prod(shape))
    return ((np.arange(n, dtype=np.int64) * mul) % mod).astype(np.float32).reshape(shape)


def make_dark_regroup(ref_models):
    """DarkCapsuleNet.forward's cell regroup (reference models.py:393-399), run through the UNMODIFIED forward: the
    backbone `model.conv` is swapped for nn.Identity() so that a synthetic [B,256,28,28] feature map reaches the
    regroup lines; a forward pre-hook on the routing layer captures what they hand to it."""
    B, g = 2, 7
    params = P()
    params.n_grid = g
    params.n_classes = 43
    params.dropout = 0.0
    model = ref_models.DarkCapsuleNet(params)
    model.conv = torch.nn.Identity()
    x = torch.from_numpy(dark_pattern((B, 256, 28, 28), 7919, 8191)).requires_grad_(True)
    seen = []
    h = model.traffic_sign_capsules.register_forward_pre_hook(lambda mod, inp: seen.append(inp[0]))
    out = model(x)
    h.remove()
    u = seen[0]                                               # [g*g*B, 512, 8]
    du = torch.from_numpy(dark_pattern(tuple(u.shape), 104729, 8179))
    u.backward(du)
    path = os.path.join(HERE, 'dark_regroup.npz')
    np.savez_compressed(path, dims=np.array([B, 256, g], dtype=np.int64), out_shape=np.array(out.shape, dtype=np.int64),
                        u=u.detach().numpy().astype(np.uint16), dx=x.grad.numpy().astype(np.uint16))
    print('%-20s x %s -> u %s -> %s (%d KB)' % ('dark_regroup', tuple(x.shape), tuple(u.shape), os.path.basename(path),
                                                os.path.getsize(path) // 1024))


def make_dark_loss(ref_loss):
    """reference darkcapsule_loss (loss_fns.py:187-204, recon off) + autograd on seeded caps / labels."""
    B, g = 4, 7
    gen = torch.Generator().manual_seed(41)
    caps = (0.6 * torch.randn(B, g, g, 5, generator=gen)).requires_grad_(True)
    y = torch.zeros(B, g, g, 48)
    y[..., 0] = (torch.rand(B, g, g, generator=gen) < 0.15).float()
    y[..., 1:5] = 0.05 + 0.9 * torch.rand(B, g, g, 4, generator=gen)
    loss = ref_loss.darkcapsule_loss(caps, y, P())
    loss.backward()
    path = os.path.join(HERE, 'dark_loss.npz')
    np.savez_compressed(path, caps=caps.detach().numpy(), y5=y[..., :5].numpy(), loss=np.float32(loss.item()),
                        dcaps=caps.grad.numpy())
    print('%-20s caps %s loss %.6f -> %s (%d KB)' % ('dark_loss', tuple(caps.shape), loss.item(), os.path.basename(path),
                                                     os.path.getsize(path) // 1024))


def make_capsnet(ref_models, ref_loss):
    """configs[0] shape: the reference's full CapsuleNet (conv1 -> primary capsules -> routing -> scores), capsule_loss
    (recon off) and autograd, batch 4, seeded construction (torch.manual_seed(0): the drop-in layers draw the same
    weights, tests/test_cabi_cpu.py) -- so the fixture only stores outputs and strided gradient probes."""
    B = 4
    params = P()
    params.n_classes = 43
    torch.manual_seed(0)
    model = ref_models.CapsuleNet(params)
    gen = torch.Generator().manual_seed(51)
    x = torch.rand(B, 3, 32, 32, generator=gen) * 2 - 1                 # U[-1,1) like utils.center_rgb
    y = torch.randint(0, 43, (B,), generator=gen)
    x.requires_grad_(True)
    scores = model(x)
    loss = ref_loss.capsule_loss(scores, y, params)
    loss.backward()
    pw = torch.cat([m.weight.grad.reshape(-1) for m in model.primary_capsules.capsules])
    pb = torch.cat([m.bias.grad.reshape(-1) for m in model.primary_capsules.capsules])
    st = 97
    path = os.path.join(HERE, 'capsnet_cfg1.npz')
    np.savez_compressed(path, dims=np.array([B, 51, st], dtype=np.int64), y=y.numpy(), scores=scores.detach().numpy(),
                        loss=np.float32(loss.item()), dx=x.grad.numpy(),
                        conv1_w_probe=model.conv1.weight.grad.reshape(-1)[::st].numpy().copy(),
                        conv1_b=model.conv1.bias.grad.numpy(), prim_w_probe=pw[::st].numpy().copy(), prim_b=pb.numpy(),
                        route_w_probe=model.traffic_sign_capsules.route_weights.grad.reshape(-1)[::st * 11].numpy().copy(),
                        x_checksum=np.float64(x.detach().double().sum().item()))
    print('%-20s scores %s loss %.6f -> %s (%d KB)' % ('capsnet_cfg1', tuple(scores.shape), loss.item(), os.path.basename(path),
                                                     os.path.getsize(path) // 1024))


def main():
    from oracle import routing_np as onp
    ref_models, ref_loss = import_reference()
    torch.manual_seed(0)
    torch.set_num_threads(1)   # one thread: reduction order (hence the fp32 bits) is reproducible
    make_primary(ref_models)
    make_dark_regroup(ref_models)
    make_dark_loss(ref_loss)
    make_capsnet(ref_models, ref_loss)
    if '--primary-only' in sys.argv:
        return
    for name, (B, N, C, K, D, R, seed, full) in CASES.items():
        u, W, y = onp.make_inputs(B, N, C, K, D, seed=seed)
        r32 = run_reference(ref_models, ref_loss, u, W, y, R, torch.float32)
        r64 = run_reference(ref_models, ref_loss, u, W, y, R, torch.float64)
        blob = dict(dims=np.array([B, N, C, K, D, R, seed], dtype=np.int64),
                    in_checksum=np.array([u.astype(np.float64).sum(), W.astype(np.float64).sum(),
                                          float(y.sum())]),
                    v=r32['v'], loss=np.float32(r32['loss']), du=r32['du'],
                    v64=r64['v'], loss64=np.float64(r64['loss']), du64=r64['du'])
        if full:
            blob.update(c=r32['c'], dW=r32['dW'], c64=r64['c'].astype(np.float64),
                        dW64_probe=r64['dW'].reshape(-1)[::7].copy(), probe_stride=np.int64(7))
        else:
            blob.update(c_probe=r32['c'].reshape(-1)[::PROBE_STRIDE].copy(),
                        c64_probe=r64['c'].reshape(-1)[::PROBE_STRIDE].copy(),
                        dW_probe=r32['dW'].reshape(-1)[::PROBE_STRIDE].copy(),
                        dW64_probe=r64['dW'].reshape(-1)[::PROBE_STRIDE].copy(),
                        dW64_sum=np.float64(r64['dW'].sum()),
                        dW64_abs_sum=np.float64(np.abs(r64['dW']).sum()),
                        probe_stride=np.int64(PROBE_STRIDE))
        path = os.path.join(HERE, name + '.npz')
        np.savez_compressed(path, **blob)
        print('%-20s B=%d N=%d C=%d K=%d D=%d R=%d loss=%.6f -> %s (%d KB)' % (
            name, B, N, C, K, D, R, r32['loss'], os.path.basename(path), os.path.getsize(path) // 1024))


if __name__ == '__main__':
    main()
