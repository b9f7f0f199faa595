import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box via gpurun)')


def golden_names():
    # routing-layer fixtures; primary_caps.npz / dark_regroup.npz (the steps before the routing layer) have their own tests
    return sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.endswith('.npz') and not f.startswith(('primary', 'dark_regroup', 'dark_loss', 'capsnet')))


def load_golden(name):
    """Returns (fixture dict, (u, W, y)) with inputs regenerated from the stored seed."""
    from oracle import routing_np as onp
    g = dict(np.load(os.path.join(GOLDEN_DIR, name + '.npz')))
    B, N, C, K, D, R, seed = [int(x) for x in g['dims']]
    u, W, y = onp.make_inputs(B, N, C, K, D, seed=seed)
    chk = np.array([u.astype(np.float64).sum(), W.astype(np.float64).sum(), float(y.sum())])
    assert np.allclose(chk, g['in_checksum'], rtol=0, atol=1e-9), 'input RNG drifted from the fixture'
    g['R'] = R
    return g, (u, W, y)


def rel_err(a, b):
    """max |a-b| / max |b|  -- the 'rel' the north_star tolerance is quoted in."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def assert_close_elementwise(a, b, rtol=1e-4, atol_frac=1e-6, what=''):
    """|a - b| <= rtol |b| + atol_frac * max|b| for EVERY element: unlike rel_err (a global max-norm), this does not let
    a large entry hide the relative error of a small one; atol_frac * max|b| is the floor below which an fp32 sum of
    that magnitude has no bits left."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    tol = rtol * np.abs(b) + atol_frac * max(np.abs(b).max(), 1e-300)
    bad = np.abs(a - b) > tol
    assert not bad.any(), '%s: %d of %d elements outside rtol=%g atol=%g*max; worst excess %.3g' % (
        what, int(bad.sum()), bad.size, rtol, atol_frac, float((np.abs(a - b) / tol).max()))


def dark_pattern(shape, mul, mod):
    """integer-valued fp32 test pattern used by tests/golden/make_golden.py::make_dark_regroup
    (exact through any permutation; inputs are regenerated from it, the fixture stores outputs only)"""
    n = int(np.prod(shape))
    return ((np.arange(n, dtype=np.int64) * mul) % mod).astype(np.float32).reshape(shape)
