"""The oracle is checked against fixtures written by the unmodified reference
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from conftest import golden_names, load_golden, rel_err
from oracle import routing_np as onp
from oracle import routing_torch as ot


@pytest.mark.parametrize('name', golden_names())
def test_numpy_oracle_fp64_matches_reference_fp64(name):
    g, (u, W, y) = load_golden(name)
    r = onp.routing_step(u.astype(np.float64), W.astype(np.float64), y, g['R'])
    # 5e-8, not 1e-12: the reference's "fp64" run still builds its logits in the default dtype
    # (models.py:72), so its iteration-0 couplings are float32(1/C) -- a 1e-8 relative wobble.
    tol = 5e-8
    assert rel_err(r['v'], g['v64']) < tol
    assert abs(r['loss'] - float(g['loss64'])) < tol
    assert rel_err(r['du'], g['du64']) < tol
    st = int(g['probe_stride'])
    assert rel_err(r['dW'].reshape(-1)[::st], g['dW64_probe']) < tol
    if 'c64' in g:
        assert rel_err(r['c'], g['c64']) < tol
    else:
        assert rel_err(r['c'].reshape(-1)[::st], g['c64_probe']) < tol


@pytest.mark.parametrize('name', golden_names())
def test_numpy_oracle_fp32_within_budget(name):
    """fp32 closed form vs the fp64 reference: this is the tolerance budget the CUDA path gets
    (north_star: rel 1e-5 on v, 1e-4 on gradients)."""
    g, (u, W, y) = load_golden(name)
    r = onp.routing_step(u, W, y, g['R'])
    assert r['v'].dtype == np.float32
    assert rel_err(r['v'], g['v64']) < 1e-5
    assert rel_err(r['du'], g['du64']) < 1e-4
    st = int(g['probe_stride'])
    assert rel_err(r['dW'].reshape(-1)[::st], g['dW64_probe']) < 1e-4
    # and the reference's own fp32 run sits inside the same budget
    assert rel_err(g['v'], g['v64']) < 1e-5
    assert rel_err(g['du'], g['du64']) < 1e-4


@pytest.mark.parametrize('name', [n for n in golden_names() if 'full' not in n])
def test_torch_port_matches_reference_fp32(name):
    """Same ATen ops in the same order: agreement to fp32 round-off (bit-identical when the
    thread count matches the one the fixture was made with)."""
    g, (u, W, y) = load_golden(name)
    torch.set_num_threads(1)
    r = ot.routing_step_t(torch.from_numpy(u), torch.from_numpy(W)[None], torch.from_numpy(y),
                          g['R'], want_c=True)
    assert rel_err(r['v'].numpy(), g['v']) < 2e-7
    assert abs(float(r['loss']) - float(g['loss'])) < 1e-6
    assert rel_err(r['du'].numpy(), g['du']) < 1e-6
    assert rel_err(r['dW'].numpy(), g['dW']) < 1e-6
    assert rel_err(r['c'].numpy(), g['c']) < 2e-7


def test_margin_loss_grad_is_gradient_of_margin_loss():
    rng = np.random.default_rng(3)
    v = onp.squash(rng.standard_normal((5, 7, 4)))
    y = rng.integers(0, 7, size=5)
    g = onp.margin_loss_grad(v, y)
    eps = 1e-6
    for idx in [(0, 0, 0), (2, 3, 1), (4, 6, 3), (1, int(y[1]), 2)]:
        vp = v.copy(); vp[idx] += eps
        vm = v.copy(); vm[idx] -= eps
        fd = (onp.margin_loss(vp, y) - onp.margin_loss(vm, y)) / (2 * eps)
        assert abs(fd - g[idx]) < 1e-7


def test_squash_bwd_matches_finite_differences():
    rng = np.random.default_rng(4)
    s = rng.standard_normal((3, 5))
    dv = rng.standard_normal((3, 5))
    ds = onp.squash_bwd(s, dv)
    eps = 1e-6
    for idx in [(0, 0), (1, 3), (2, 4)]:
        sp = s.copy(); sp[idx] += eps
        sm = s.copy(); sm[idx] -= eps
        fd = ((onp.squash(sp) - onp.squash(sm)) * dv).sum() / (2 * eps)
        assert abs(fd - ds[idx]) < 1e-7


def test_single_class_capsule_ignores_iterations():
    """with one class capsule the softmax is identically 1."""
    u, W, _ = onp.make_inputs(3, 64, 1, 8, 5, seed=5, dtype=np.float64)
    assert np.array_equal(onp.routing_forward(u, W, 1), onp.routing_forward(u, W, 3))


def test_zero_input_gives_nan_like_the_reference():
    """squash has no epsilon (models.py:64-67): a zero capsule sum is 0/0."""
    u = np.zeros((1, 4, 8)); W = np.ones((4, 3, 8, 2))
    with np.errstate(invalid='ignore'):
        assert np.isnan(onp.routing_forward(u, W, 2)).all()


@pytest.mark.parametrize('name', golden_names())
def test_c_oracle_matches_reference(name):
    from oracle import routing_c as oc
    g, (u, W, y) = load_golden(name)
    st = int(g['probe_stride'])
    r = oc.routing_step(u.astype(np.float64), W.astype(np.float64), y, g['R'])
    assert rel_err(r['v'], g['v64']) < 5e-8
    assert abs(r['loss'] - float(g['loss64'])) < 5e-8
    assert rel_err(r['du'], g['du64']) < 5e-8
    assert rel_err(r['dW'].reshape(-1)[::st], g['dW64_probe']) < 5e-8
    r = oc.routing_step(u, W, y, g['R'])
    assert rel_err(r['v'], g['v64']) < 1e-5
    assert rel_err(r['du'], g['du64']) < 1e-4
    assert rel_err(r['dW'].reshape(-1)[::st], g['dW64_probe']) < 1e-4


def test_primary_tail_matches_reference():
    """The primary-capsule tail (reference models.py:81-82: K views + cat + squash) and its gradient,
    against the reference layer's own output / autograd (tests/golden/primary_caps.npz)."""
    import os
    from conftest import GOLDEN_DIR
    from oracle import routing_np as onp
    g = dict(np.load(os.path.join(GOLDEN_DIR, 'primary_caps.npz')))
    K = int(g['dims'][3])
    u = onp.primary_tail(g['conv'].astype(np.float64), K)
    assert u.shape == g['u'].shape
    assert rel_err(u, g['u']) < 1e-6
    dconv = onp.primary_tail_bwd(g['conv'].astype(np.float64), g['du'].astype(np.float64), K)
    assert rel_err(dconv, g['dconv']) < 1e-6


def test_dark_regroup_matches_reference():
    """DarkCapsuleNet's cell regroup (reference models.py:393-399) and its gradient: bit-exact against
    what the unmodified DarkCapsuleNet.forward handed to its routing layer (tests/golden/dark_regroup.npz)."""
    import os
    from conftest import GOLDEN_DIR, dark_pattern
    from oracle import routing_np as onp
    g = dict(np.load(os.path.join(GOLDEN_DIR, 'dark_regroup.npz')))
    B, Cch, grid = [int(v) for v in g['dims']]
    x = dark_pattern((B, Cch, 4 * grid, 4 * grid), 7919, 8191)
    u = onp.dark_regroup(x, grid)
    assert np.array_equal(u, g['u'].astype(np.float32))
    du = dark_pattern(u.shape, 104729, 8179)
    dx = onp.dark_regroup_bwd(du, B, Cch, grid)
    assert np.array_equal(dx.reshape(x.shape), g['dx'].astype(np.float32))


def test_dark_loss_matches_reference():
    """darkcapsule_loss + polar_transform (reference loss_fns.py:187-204, utils.py:69-85) and the gradient
    autograd gives for it (tests/golden/dark_loss.npz)."""
    import os
    from conftest import GOLDEN_DIR
    from oracle import routing_np as onp
    g = dict(np.load(os.path.join(GOLDEN_DIR, 'dark_loss.npz')))
    y = np.zeros(g['y5'].shape[:3] + (48,))
    y[..., :5] = g['y5']
    loss, dcaps = onp.dark_loss(g['caps'].astype(np.float64), y)
    assert abs(loss - float(g['loss'])) < 1e-6 * abs(float(g['loss']))
    assert rel_err(dcaps, g['dcaps']) < 1e-6
