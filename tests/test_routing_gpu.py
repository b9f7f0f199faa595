"""Parity of the CUDA path (through the C ABI / the CapsuleLayer module) against the oracle and
the reference-generated golden fixtures.  Needs a B200: run with `pytest -m gpu`.

Tolerances are the north_star's: rel 1e-5 on v (and c, loss), rel 1e-4 on gradients, where
rel = max|a-b| / max|b| and b is the reference's fp64 result (fp32 result for the fixtures' fp32
copies, which differ from fp64 by up to ~1e-6 themselves)."""
import numpy as np
import pytest
import torch

from conftest import assert_close_elementwise, golden_names, load_golden, rel_err

pytestmark = pytest.mark.gpu

TOL_V = 1e-5
TOL_G = 1e-4
KNOB_DEFAULTS = {'c1v': 2, 'spt': 0, 'isplit': 0, 'tc': 1, 'fused': 1, 'fsws': 1, 'c1': 1, 'gradmma': 1, 'gradjw': 0, 'hostmb': 0, 'sbstaged': 1, 'tcstages': 10}


@pytest.fixture(autouse=True)
def _reset_tuning_knobs():
    """The knobs are process-global: a test that fails between set and reset must not leak its setting into the next."""
    yield
    from cs231_capsule_yolo_traffic_sign_detection_b200 import _cabi
    for k, v in KNOB_DEFAULTS.items():
        _cabi.set_tuning(k, v)


@pytest.fixture(scope='module')
def capsb():
    import cs231_capsule_yolo_traffic_sign_detection_b200 as m
    assert torch.cuda.is_available(), 'gpu tests need a CUDA device'
    m._cabi.lib()                       # raises if the CUDA library is missing: no silent fallback
    return m


def cuda_step(capsb, u, W, y, R, want_c=True, grad_v_extra=None):
    """fwd (+c) + fused margin loss + bwd through the public API; returns numpy dict."""
    dev = torch.device('cuda')
    ut = torch.from_numpy(u).to(dev).requires_grad_(True)
    Wt = torch.from_numpy(W)[None].to(dev).requires_grad_(True)
    yt = torch.from_numpy(y).to(dev)
    out = {}
    if want_c:
        with torch.no_grad():
            _, c = capsb.dynamic_routing(ut, Wt, R, return_couplings=True)
        out['c'] = c.cpu().numpy()
    v, loss = capsb.routing_margin_loss(ut, Wt, yt, R)
    total = loss
    if grad_v_extra is not None:
        total = loss + (v * torch.from_numpy(grad_v_extra).to(dev)).sum()
    total.backward()
    torch.cuda.synchronize()
    out.update(v=v.detach().cpu().numpy(), loss=float(loss), du=ut.grad.cpu().numpy(),
               dW=Wt.grad[0].cpu().numpy())
    return out


@pytest.mark.parametrize('name', golden_names())
def test_golden_fixture(capsb, name):
    g, (u, W, y) = load_golden(name)
    r = cuda_step(capsb, u, W, y, g['R'])
    st = int(g['probe_stride'])
    assert rel_err(r['v'], g['v64']) < TOL_V
    assert abs(r['loss'] - float(g['loss64'])) < TOL_V * max(1.0, abs(float(g['loss64'])))
    assert rel_err(r['du'], g['du64']) < TOL_G
    assert rel_err(r['dW'].reshape(-1)[::st], g['dW64_probe']) < TOL_G
    if 'c64' in g:
        assert rel_err(r['c'], g['c64']) < TOL_V
        assert rel_err(r['dW'], g['dW']) < TOL_G          # full dW against the reference's fp32 run
    else:
        assert rel_err(r['c'].reshape(-1)[::st], g['c64_probe']) < TOL_V
    assert np.allclose(r['c'].sum(-1), 1.0, atol=1e-5)    # couplings are a distribution over j


@pytest.mark.parametrize('dims', [
    (37, 1296, 43, 8, 16, 3),     # CapsuleNet routing shape, ragged lane tile
    (64, 1152, 43, 8, 16, 3),     # BASELINE.json config 2 shape
    (100, 130, 43, 8, 16, 2),
    (70, 512, 1, 8, 5, 3),        # DarkCapsuleNet head
    (33, 64, 10, 8, 16, 4),
    (5, 50, 43, 8, 21, 3),        # DarkCapsuleNet3 head dims
    (6, 40, 49, 8, 48, 3),        # DarkCapsuleNet2 head dims
    (40, 96, 43, 8, 32, 5),       # sweep corner: D=32, R=5
    (3, 7, 3, 8, 4, 3),           # tiny / odd everything
    (2, 9, 2, 8, 24, 2),
    (70, 200, 43, 8, 24, 3),      # sweep corner: D=24 (tcgen05 passes with N = 96, FMA gradient sweep)
    (130, 64, 5, 8, 32, 2),       # D=32, two j-groups, second one ragged
    (45, 72, 10, 8, 32, 3),       # D=32 on the mma gradient kernel (C >= 7): 20 pseudo-capsules, two CTA rows
    (50, 40, 9, 8, 12, 3),        # D=12 padded to 16 on the tensor-core kernels
    (36, 48, 8, 8, 40, 2),        # D=40 padded to 48: three pseudo-capsules, the last one half empty
])
def test_against_c_oracle_fp64(capsb, dims):
    from oracle import routing_c as oc
    from oracle import routing_np as onp
    B, N, C, K, D, R = dims
    u, W, y = onp.make_inputs(B, N, C, K, D, seed=sum(dims))
    ref = oc.routing_step(u.astype(np.float64), W.astype(np.float64), y, R)
    r = cuda_step(capsb, u, W, y, R)
    assert rel_err(r['v'], ref['v']) < TOL_V
    assert rel_err(r['c'], ref['c']) < TOL_V
    assert abs(r['loss'] - ref['loss']) < TOL_V * max(1.0, abs(ref['loss']))
    assert rel_err(r['du'], ref['du']) < TOL_G
    assert rel_err(r['dW'], ref['dW']) < TOL_G


def test_external_grad_v_and_unfused_loss(capsb):
    """Gradient arriving through autograd (decoder / coordinate losses feed v directly:
    reference models.py:122, loss_fns.py:197) plus the unfused margin loss written with torch ops
    like reference loss_fns.py:11-23 must match the fused kernel path and the oracle."""
    from oracle import routing_np as onp
    B, N, C, K, D, R = 19, 80, 43, 8, 16, 3
    u, W, y = onp.make_inputs(B, N, C, K, D, seed=5)
    gext = np.random.default_rng(6).standard_normal((B, C, D)).astype(np.float32) * 0.01
    ref = onp.routing_step(u.astype(np.float64), W.astype(np.float64), y, R, grad_v_extra=gext.astype(np.float64))
    fused = cuda_step(capsb, u, W, y, R, want_c=False, grad_v_extra=gext)
    assert rel_err(fused['du'], ref['du']) < TOL_G
    assert rel_err(fused['dW'], ref['dW']) < TOL_G

    dev = torch.device('cuda')
    ut = torch.from_numpy(u).to(dev).requires_grad_(True)
    Wt = torch.from_numpy(W)[None].to(dev).requires_grad_(True)
    v = capsb.dynamic_routing(ut, Wt, R)
    scores = (v ** 2).sum(dim=-1) ** 0.5
    onehot = torch.eye(C, device=dev).index_select(0, torch.from_numpy(y).to(dev))
    margin = onehot * torch.relu(0.9 - scores) ** 2 + 0.5 * (1 - onehot) * torch.relu(scores - 0.1) ** 2
    loss = margin.sum() / B + (v * torch.from_numpy(gext).to(dev)).sum()
    loss.backward()
    assert rel_err(ut.grad.cpu().numpy(), ref['du']) < TOL_G
    assert rel_err(Wt.grad[0].cpu().numpy(), ref['dW']) < TOL_G
    assert abs(float(margin.sum() / B) - ref['loss']) < 1e-5


def test_bit_reproducible_and_tuning_invariant(capsb):
    """Fixed summation orders: same bits run to run; samples-per-thread / i-split only regroup
    independent work (v is bit-identical per sample when the i-split is unchanged)."""
    from oracle import routing_np as onp
    B, N, C, K, D, R = 150, 200, 43, 8, 16, 3
    u, W, y = onp.make_inputs(B, N, C, K, D, seed=9)
    base = cuda_step(capsb, u, W, y, R, want_c=False)
    again = cuda_step(capsb, u, W, y, R, want_c=False)
    for k in ('v', 'du', 'dW'):
        assert np.array_equal(base[k], again[k]), k
    try:
        for spt in (1, 2, 4):
            capsb._cabi.set_tuning('spt', spt)
            capsb._cabi.set_tuning('isplit', 5)
            r = cuda_step(capsb, u, W, y, R, want_c=False)
            if spt == 1:
                first = r
            assert np.array_equal(r['v'], first['v'])
            assert np.array_equal(r['du'], first['du'])
            assert rel_err(r['dW'], base['dW']) < 1e-5
        capsb._cabi.set_tuning('spt', 0)
        for isplit in (1, 3, 32):
            capsb._cabi.set_tuning('isplit', isplit)
            r = cuda_step(capsb, u, W, y, R, want_c=False)
            assert rel_err(r['v'], base['v']) < 1e-5
            assert rel_err(r['dW'], base['dW']) < 1e-5
    finally:
        capsb._cabi.set_tuning('spt', 0)
        capsb._cabi.set_tuning('isplit', 0)


def test_tensor_core_and_fma_engines_agree(capsb):
    """The tcgen05 pass kernel / mma.sync gradient kernel (3xTF32) and the plain fp32-FMA kernels are
    two implementations of the same contract; both must sit inside the oracle's tolerance and
    agree with each other to fp32 round-off."""
    from oracle import routing_c as oc
    from oracle import routing_np as onp
    B, N, C, K, D, R = 200, 160, 43, 8, 16, 3
    u, W, y = onp.make_inputs(B, N, C, K, D, seed=21)
    ref = oc.routing_step(u.astype(np.float64), W.astype(np.float64), y, R)
    res = {}
    try:
        for name, tc, gm in (('tensor', 1, 1), ('fma', 0, 0), ('mixed', 1, 0)):
            capsb._cabi.set_tuning('tc', tc)
            capsb._cabi.set_tuning('gradmma', gm)
            res[name] = cuda_step(capsb, u, W, y, R)
            for k, tol in (('v', TOL_V), ('c', TOL_V), ('du', TOL_G), ('dW', TOL_G)):
                assert rel_err(res[name][k], ref[k]) < tol, (name, k)
    finally:
        capsb._cabi.set_tuning('tc', 1)
        capsb._cabi.set_tuning('gradmma', 1)
    for k in ('v', 'c', 'du', 'dW'):
        assert rel_err(res['tensor'][k], res['fma'][k]) < 5e-6, k
    # 8 instead of 11 capsules per CTA in the gradient kernel: dW is bit-identical (the batch sum of one (i, j)
    # never crosses warps), du only regroups the sum over j
    try:
        capsb._cabi.set_tuning('gradjw', 8)
        alt = cuda_step(capsb, u, W, y, R)
    finally:
        capsb._cabi.set_tuning('gradjw', 0)
    assert np.array_equal(alt['dW'], res['tensor']['dW'])
    assert rel_err(alt['du'], res['tensor']['du']) < 5e-6


@pytest.mark.parametrize('dims', [
    (200, 160, 43, 8, 16, 3),     # 6 capsule groups (cluster of 6), ragged sample quad, last group 3 capsules wide
    (129, 97, 10, 8, 16, 4),      # cluster of 2, N not a multiple of the split granule
    (64, 64, 5, 8, 16, 2),        # one group: cluster of 1
    (300, 72, 64, 8, 16, 3),      # cluster of 8 (the portable maximum)
    (40, 256, 43, 8, 12, 5),      # D = 12 padded to 16, R = 5
    (31, 40, 65, 8, 16, 3),       # 9 groups: beyond the cluster limit -> must take the unfused path by itself
])
def test_fused_sweep_matches_unfused_and_oracle(capsb, dims):
    """The cluster-fused sweep (logits -> softmax -> weighted sum in one kernel, partial normalisers exchanged through
    distributed shared memory) and the three-kernel path are two implementations of reference models.py:75-79 and of
    its backward: both inside the oracle's tolerance, and within fp32 round-off of each other."""
    from oracle import routing_c as oc
    from oracle import routing_np as onp
    B, N, C, K, D, R = dims
    u, W, y = onp.make_inputs(B, N, C, K, D, seed=sum(dims) + 1)
    ref = oc.routing_step(u.astype(np.float64), W.astype(np.float64), y, R)
    res = {}
    for fused in (1, 0):
        capsb._cabi.set_tuning('fused', fused)
        res[fused] = cuda_step(capsb, u, W, y, R)
        for k, tol in (('v', TOL_V), ('c', TOL_V), ('du', TOL_G), ('dW', TOL_G)):
            assert rel_err(res[fused][k], ref[k]) < tol, (fused, k)
        assert abs(res[fused]['loss'] - ref['loss']) < TOL_V * max(1.0, abs(ref['loss']))
        assert_close_elementwise(res[fused]['dW'], ref['dW'], what='dW fused=%d' % fused)
        assert_close_elementwise(res[fused]['du'], ref['du'], what='du fused=%d' % fused)
    for k in ('v', 'c', 'du', 'dW'):
        assert rel_err(res[1][k], res[0][k]) < 5e-6, k
    # the two epilogue organisations of the fused sweep (8 logit + 8 accumulate warps / 8 warps doing both) run the
    # same arithmetic in the same order
    capsb._cabi.set_tuning('fused', 1)
    capsb._cabi.set_tuning('fsws', 0)
    old_epi = cuda_step(capsb, u, W, y, R)
    capsb._cabi.set_tuning('fsws', 1)
    for k in ('v', 'c', 'du', 'dW'):
        assert rel_err(old_epi[k], res[1][k]) < 5e-6, k
    # i-split of the fused sweep only regroups the sum over input capsules
    capsb._cabi.set_tuning('isplit', 2)
    alt = cuda_step(capsb, u, W, y, R)
    for k in ('v', 'du', 'dW'):
        assert rel_err(alt[k], res[1][k]) < 5e-6, k


@pytest.mark.parametrize('B', [512])
def test_benchmark_shape_against_oracle(capsb, B):
    """The configuration bench.py reports (BASELINE.json configs[1]: 1152 -> 43 x 16, 3 iterations) at a batch of
    several 128-sample tiles, every output -- v, c, du AND dW -- against the fp64 C oracle, max-norm and per element."""
    from oracle import routing_c as oc
    from oracle import routing_np as onp
    N, C, K, D, R = 1152, 43, 8, 16, 3
    u, W, y = onp.make_inputs(B, N, C, K, D, seed=77)
    ref = oc.routing_step(u.astype(np.float64), W.astype(np.float64), y, R)
    r = cuda_step(capsb, u, W, y, R)
    assert rel_err(r['v'], ref['v']) < TOL_V
    assert rel_err(r['c'], ref['c']) < TOL_V
    assert abs(r['loss'] - ref['loss']) < TOL_V * max(1.0, abs(ref['loss']))
    assert rel_err(r['du'], ref['du']) < TOL_G
    assert rel_err(r['dW'], ref['dW']) < TOL_G
    assert_close_elementwise(r['v'], ref['v'], rtol=1e-5, atol_frac=1e-6, what='v')
    assert_close_elementwise(r['du'], ref['du'], what='du')
    assert_close_elementwise(r['dW'], ref['dW'], what='dW')


def test_benchmark_shape_host_step_against_oracle(capsb):
    """The end-to-end call bench.py times (caps_route_step_host, three pipelined micro-batches from B >= 2048) at the
    benchmark shape, against the fp64 C oracle: loss, v, du and the micro-batch-summed dW."""
    from oracle import routing_c as oc
    from oracle import routing_np as onp
    B, N, C, K, D, R = 2048 + 128, 1152, 43, 8, 16, 3
    u, W, y = onp.make_inputs(B, N, C, K, D, seed=78)
    ref = oc.routing_step(u.astype(np.float64), W.astype(np.float64), y, R, want_c=False)
    dev = torch.device('cuda')
    step = capsb.HostStep(B, N, C, K, D, R)
    Wd = torch.from_numpy(W).to(dev)
    dWd = torch.empty_like(Wd)
    uh, yh = torch.from_numpy(u).pin_memory(), torch.from_numpy(y).pin_memory()
    vh, duh = torch.empty(B, C, D).pin_memory(), torch.empty(B, N, K).pin_memory()
    loss = float(step(uh, yh, Wd, dWd, v_host=vh, du_host=duh)[0])
    assert abs(loss - ref['loss']) < TOL_V * max(1.0, abs(ref['loss']))
    assert rel_err(vh.numpy(), ref['v']) < TOL_V
    assert rel_err(duh.numpy(), ref['du']) < TOL_G
    assert rel_err(dWd.cpu().numpy(), ref['dW']) < TOL_G
    assert_close_elementwise(dWd.cpu().numpy(), ref['dW'], what='dW')
    assert_close_elementwise(duh.numpy(), ref['du'], what='du')


def test_backward_validates_forward_state(capsb):
    """caps_route_backward on a workspace no forward filled, filled for other dims, or filled without with_grad
    returns CAPS_E_STATE (-4) instead of reading garbage; and it replays the engines the forward used even if the
    tuning knobs changed in between."""
    from oracle import routing_np as onp
    from cs231_capsule_yolo_traffic_sign_detection_b200 import _cabi
    L = _cabi.lib()
    dev = torch.device('cuda')
    B, N, C, K, D, R = 40, 64, 43, 8, 16, 3
    u, W, y = onp.make_inputs(B, N, C, K, D, seed=4)
    ut, Wt, yt = torch.from_numpy(u).to(dev), torch.from_numpy(W).to(dev), torch.from_numpy(y).to(dev)
    nbytes = L.caps_route_workspace_bytes(B, N, C, K, D, R, 1)
    st = torch.cuda.current_stream().cuda_stream
    P = lambda t: t.data_ptr()
    v, du, dW = torch.empty(B, C, D, device=dev), torch.empty(B, N, K, device=dev), torch.empty_like(Wt)

    def bwd(ws, b=B):
        return L.caps_route_backward(P(ut), P(Wt), None, P(yt), 1.0 / B, None, P(du), P(dW), P(ws), ws.numel(), b, N, C, K, D, R, st)
    fresh = torch.empty(nbytes + 4096, dtype=torch.uint8, device=dev)
    assert bwd(fresh) == -4 and b'no caps_route_forward' in L.caps_last_error()
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    assert L.caps_route_forward(P(ut), P(Wt), P(v), None, P(ws), nbytes, B, N, C, K, D, R, 0, st) == 0
    assert bwd(ws) == -4                                   # forward ran without with_grad
    assert L.caps_route_forward(P(ut), P(Wt), P(v), None, P(ws), nbytes, B, N, C, K, D, R, 1, st) == 0
    assert bwd(ws, b=B - 8) == -4                          # other dims
    # knobs flipped between forward and backward: the backward must still use what the forward prepared
    _cabi.set_tuning('tc', 0)
    _cabi.set_tuning('fused', 0)
    assert bwd(ws) == 0
    torch.cuda.synchronize()
    ref = cuda_step(capsb, u, W, y, R, want_c=False)       # knobs at defaults again?  no: still flipped -> FMA engines
    _cabi.set_tuning('tc', 1)
    _cabi.set_tuning('fused', 1)
    assert rel_err(dW.cpu().numpy(), ref['dW']) < 5e-6
    assert rel_err(du.cpu().numpy(), ref['du']) < 5e-6


@pytest.mark.parametrize('dims', [
    (1568, 512, 1, 8, 5, 3),      # the DarkCapsuleNet head at its real routing batch (32 images x 49 cells)
    (7, 512, 1, 8, 5, 1),         # odd batch (the forward kernel takes samples in pairs), one iteration
    (33, 24, 1, 8, 8, 2),         # D = 8, N*8 smaller than one pass of the thread block
    (20, 100, 1, 8, 3, 3),
    (4100, 12, 1, 8, 5, 2),       # more samples than one shared-memory chunk of ds in the backward kernel (4096)
])
def test_single_capsule_head_kernels(capsb, dims):
    """One class capsule (reference models.py:368-370): the dedicated GEMM + squash kernels (caps_c1.cu) against the
    fp64 oracle and against the general kernels (tuning knob c1 = 0), with the fused margin gradient and an external
    grad_v, as DarkCapsuleNet's loss feeds it (loss_fns.py:197)."""
    from oracle import routing_c as oc
    from oracle import routing_np as onp
    B, N, C, K, D, R = dims
    u, W, y = onp.make_inputs(B, N, C, K, D, seed=sum(dims))
    gext = (np.random.default_rng(3).standard_normal((B, C, D)) * 0.05).astype(np.float32)
    ref = oc.routing_step(u.astype(np.float64), W.astype(np.float64), y, R, grad_v_extra=gext.astype(np.float64))
    res = {}
    for c1 in (2, 1, 0):                # 2 / 1: the two generations of the dedicated kernels (knob c1v), 0: the general kernels
        capsb._cabi.set_tuning('c1', 1 if c1 else 0)
        capsb._cabi.set_tuning('c1v', c1 if c1 else 2)
        res[c1] = cuda_step(capsb, u, W, y, R, grad_v_extra=gext)
        assert rel_err(res[c1]['v'], ref['v']) < TOL_V, c1
        assert np.array_equal(res[c1]['c'], np.ones((B, N, 1), np.float32))
        assert abs(res[c1]['loss'] - ref['loss']) < TOL_V * max(1.0, abs(ref['loss']))
        assert rel_err(res[c1]['du'], ref['du']) < TOL_G, c1
        assert rel_err(res[c1]['dW'], ref['dW']) < TOL_G, c1
        assert_close_elementwise(res[c1]['dW'], ref['dW'], what='dW c1=%d' % c1)
    for k in ('v', 'du', 'dW'):
        assert rel_err(res[1][k], res[0][k]) < 5e-6, k
        assert rel_err(res[2][k], res[0][k]) < 5e-6, k
    capsb._cabi.set_tuning('c1', 1)
    twice = cuda_step(capsb, u, W, y, R, grad_v_extra=gext)         # second generation again: fixed summation order
    for k in ('v', 'du', 'dW'):
        assert np.array_equal(twice[k], res[2][k]), k
    capsb._cabi.set_tuning('c1', 0)
    again = cuda_step(capsb, u, W, y, R, grad_v_extra=gext)          # c1 = 0: fixed summation order either way
    for k in ('v', 'du', 'dW'):
        assert np.array_equal(again[k], res[0][k]), k


def test_batch_permutation_and_additivity_at_full_size(capsb):
    """Size-independent properties at BASELINE.json's shape (1152 -> 43x16, R=3), where the
    oracle is too slow: samples are independent, so (a) permuting the batch permutes v and du
    bit-exactly, (b) dW of a batch is the sum of the dW of its halves (fp32 reassociation only),
    (c) a sample's result does not depend on its batch-mates."""
    from oracle import routing_np as onp
    from oracle import routing_c as oc
    B, N, C, K, D, R = 1024, 1152, 43, 8, 16, 3
    u, W, y = onp.make_inputs(B, N, C, K, D, seed=3)
    dev = torch.device('cuda')
    g = torch.Generator(device='cpu').manual_seed(1)
    gext = (torch.randn(B, C, D, generator=g) * 0.01).numpy()

    def run(idx):
        ut = torch.from_numpy(u[idx]).to(dev).requires_grad_(True)
        Wt = torch.from_numpy(W)[None].to(dev).requires_grad_(True)
        v = capsb.dynamic_routing(ut, Wt, R)
        (v * torch.from_numpy(gext[idx]).to(dev)).sum().backward()
        return v.detach().cpu().numpy(), ut.grad.cpu().numpy(), Wt.grad[0].cpu().numpy()

    ident = np.arange(B)
    perm = np.random.default_rng(0).permutation(B)
    capsb._cabi.set_tuning('isplit', 4)      # same split of the i range for every batch size below
    v0, du0, dW0 = run(ident)
    v1, du1, dW1 = run(perm)
    assert np.array_equal(v1, v0[perm])
    assert np.array_equal(du1, du0[perm])
    assert rel_err(dW1, dW0) < 1e-5
    va, dua, dWa = run(ident[:500])
    vb, dub, dWb = run(ident[500:])
    assert np.array_equal(va, v0[:500]) and np.array_equal(vb, v0[500:])
    assert rel_err(dWa + dWb, dW0) < 1e-5
    # spot-check a few samples of the big batch against the fp64 oracle
    pick = np.array([0, 499, 500, 1023])
    ref = oc.routing_step(u[pick].astype(np.float64), W.astype(np.float64), None, R,
                          grad_v_extra=gext[pick].astype(np.float64))
    capsb._cabi.set_tuning('isplit', 0)
    assert rel_err(v0[pick], ref['v']) < TOL_V
    assert rel_err(du0[pick], ref['du']) < TOL_G


def test_edge_cases(capsb):
    from oracle import routing_np as onp
    dev = torch.device('cuda')
    # empty batch
    W = torch.randn(1, 12, 5, 8, 16, device=dev, requires_grad=True)
    u = torch.zeros(0, 12, 8, device=dev, requires_grad=True)
    v = capsb.dynamic_routing(u, W, 3)
    assert v.shape == (0, 5, 16)
    v.sum().backward()
    assert torch.count_nonzero(W.grad) == 0
    # a zero capsule sum is 0/0 in the reference's squash (models.py:64-67): NaN, not clamped
    v = capsb.dynamic_routing(torch.zeros(2, 12, 8, device=dev), W.detach(), 3)
    assert torch.isnan(v).all()
    # single class capsule: routing iterations are no-ops (SURVEY 3.3 iv)
    u, Wn, _ = onp.make_inputs(4, 64, 1, 8, 5, seed=2)
    a = capsb.dynamic_routing(torch.from_numpy(u).to(dev), torch.from_numpy(Wn)[None].to(dev), 1)
    b = capsb.dynamic_routing(torch.from_numpy(u).to(dev), torch.from_numpy(Wn)[None].to(dev), 3)
    assert torch.equal(a, b)
    # unsupported shapes fail loudly
    with pytest.raises(RuntimeError):
        capsb.dynamic_routing(torch.zeros(2, 4, 6, device=dev), torch.zeros(1, 4, 3, 6, 16, device=dev), 3)
    with pytest.raises(RuntimeError):
        capsb.dynamic_routing(torch.zeros(2, 4, 8), torch.zeros(1, 4, 3, 8, 16), 3)   # CPU tensors


def test_host_step_matches_device_path(capsb):
    """The HOST-buffer entry point (what bench.py's e2e leg times) gives the device path's bits."""
    from oracle import routing_np as onp
    B, N, C, K, D, R = 48, 96, 43, 8, 16, 3
    u, W, y = onp.make_inputs(B, N, C, K, D, seed=8)
    ref = cuda_step(capsb, u, W, y, R, want_c=False)
    dev = torch.device('cuda')
    step = capsb.HostStep(B, N, C, K, D, R)
    Wd = torch.from_numpy(W).to(dev)
    dWd = torch.empty_like(Wd)
    uh = torch.from_numpy(u).pin_memory()
    yh = torch.from_numpy(y).pin_memory()
    vh = torch.empty(B, C, D).pin_memory()
    duh = torch.empty(B, N, K).pin_memory()
    loss = step(uh, yh, Wd, dWd, v_host=vh, du_host=duh)
    assert abs(float(loss[0]) - ref['loss']) < 1e-7
    assert np.array_equal(vh.numpy(), ref['v'])
    assert np.array_equal(duh.numpy(), ref['du'])
    assert np.array_equal(dWd.cpu().numpy(), ref['dW'])


def test_host_step_micro_batches(capsb):
    """From B >= 2048 the host step pipelines three micro-batches (copy under compute); v / du are
    per-sample so they keep the device path's bits, dW and the loss are re-associated sums."""
    from oracle import routing_np as onp
    from cs231_capsule_yolo_traffic_sign_detection_b200 import _cabi
    B, N, C, K, D, R = 2048 + 128, 40, 10, 8, 16, 3
    u, W, y = onp.make_inputs(B, N, C, K, D, seed=9)
    dev = torch.device('cuda')
    Wd = torch.from_numpy(W).to(dev)
    uh = torch.from_numpy(u).pin_memory()
    yh = torch.from_numpy(y).pin_memory()
    outs = []
    for hostmb in (1, 0):
        _cabi.set_tuning('hostmb', hostmb)
        _cabi.set_tuning('isplit', 4)
        try:
            step = capsb.HostStep(B, N, C, K, D, R)
            dWd = torch.empty_like(Wd)
            vh = torch.empty(B, C, D).pin_memory()
            duh = torch.empty(B, N, K).pin_memory()
            loss = float(step(uh, yh, Wd, dWd, v_host=vh, du_host=duh)[0])
            outs.append((loss, vh.numpy().copy(), duh.numpy().copy(), dWd.cpu().numpy()))
            if hostmb == 0:      # twice the same call: same bits (fixed summation order over micro-batches)
                dW4 = torch.empty_like(Wd)
                step(uh, yh, Wd, dW4)
                assert np.array_equal(dW4.cpu().numpy(), outs[-1][3])
        finally:
            _cabi.set_tuning('hostmb', 0)
            _cabi.set_tuning('isplit', 0)
    (l1, v1, du1, dW1), (l3, v3, du3, dW3) = outs
    assert abs(l1 - l3) < 1e-6 * max(1.0, abs(l1))
    assert np.array_equal(v1, v3)
    assert np.array_equal(du1, du3)
    assert rel_err(dW3, dW1) < 1e-5


def test_squash_kernel(capsb):
    from oracle import routing_np as onp
    dev = torch.device('cuda')
    x = torch.randn(1000, 8, device=dev, requires_grad=True)
    layer = capsb.CapsuleLayer(None, n_caps=3, n_nodes=4, in_C=8, out_C=16)
    y = layer.squash(x)
    ref = onp.squash(x.detach().cpu().numpy().astype(np.float64))
    assert rel_err(y.detach().cpu().numpy(), ref) < 1e-6
    gy = torch.randn_like(y)
    y.backward(gy)
    gref = onp.squash_bwd(x.detach().cpu().numpy().astype(np.float64), gy.cpu().numpy().astype(np.float64))
    assert rel_err(x.grad.cpu().numpy(), gref) < 1e-5


def test_primary_capsule_tail_kernel(capsb):
    """caps_primary_squash / _backward (the K views + cat + squash of reference models.py:81-82 in one
    pass) against the reference's own output and autograd gradient, and against the oracle at the
    real CapsuleNet shape (8 capsules x 16 channels x 9x9)."""
    import os
    from conftest import GOLDEN_DIR
    from oracle import routing_np as onp
    from cs231_capsule_yolo_traffic_sign_detection_b200.capsule import _PrimarySquashFn
    dev = torch.device('cuda')
    g = dict(np.load(os.path.join(GOLDEN_DIR, 'primary_caps.npz')))
    K = int(g['dims'][3])
    conv = torch.from_numpy(g['conv']).to(dev).requires_grad_(True)
    u = _PrimarySquashFn.apply(conv, K)
    assert rel_err(u.detach().cpu().numpy(), g['u']) < 1e-6
    u.backward(torch.from_numpy(g['du']).to(dev))
    assert rel_err(conv.grad.cpu().numpy(), g['dconv']) < 1e-5
    rng = np.random.default_rng(5)
    for (B, Kc, Cc, H) in ((33, 8, 16, 9), (2, 3, 7, 5), (4, 12, 4, 3)):
        cv = rng.standard_normal((B, Kc * Cc, H, H)).astype(np.float32)
        du = rng.standard_normal((B, Cc * H * H, Kc)).astype(np.float32)
        ct = torch.from_numpy(cv).to(dev).requires_grad_(True)
        ut = _PrimarySquashFn.apply(ct, Kc)
        assert rel_err(ut.detach().cpu().numpy(), onp.primary_tail(cv.astype(np.float64), Kc)) < 1e-6
        ut.backward(torch.from_numpy(du).to(dev))
        assert rel_err(ct.grad.cpu().numpy(), onp.primary_tail_bwd(cv.astype(np.float64), du.astype(np.float64), Kc)) < 1e-5


def test_primary_capsule_layer_matches_reference(capsb):
    """The conv->caps branch of the drop-in layer (ONE convolution over the concatenated weights + the
    fused tail) against the reference layer's output and autograd gradients; the parameters stay
    K separate Conv2d modules (state_dict names and shapes of the reference)."""
    import os
    from conftest import GOLDEN_DIR
    dev = torch.device('cuda')
    g = dict(np.load(os.path.join(GOLDEN_DIR, 'primary_caps.npz')))
    B, Cin, H, K, Cc, kern, stride = [int(v) for v in g['dims']]
    layer = capsb.CapsuleLayer(None, n_caps=K, n_nodes=-1, in_C=Cin, out_C=Cc, kernel=kern, stride=stride).to(dev)
    assert sorted(layer.state_dict().keys()) == sorted(
        ['capsules.%d.%s' % (k, n) for k in range(K) for n in ('weight', 'bias')])
    with torch.no_grad():
        for k, m in enumerate(layer.capsules):
            m.weight.copy_(torch.from_numpy(g['weight'][k]))
            m.bias.copy_(torch.from_numpy(g['bias'][k]))
    torch.backends.cudnn.allow_tf32 = False          # fp32 convolution like the reference's
    x = torch.from_numpy(g['x']).to(dev).requires_grad_(True)
    u = layer(x)
    assert tuple(u.shape) == tuple(g['u'].shape)
    assert rel_err(u.detach().cpu().numpy(), g['u']) < 1e-5
    u.backward(torch.from_numpy(g['du']).to(dev))
    assert rel_err(x.grad.cpu().numpy(), g['dx']) < 1e-4
    dw = np.stack([m.weight.grad.cpu().numpy() for m in layer.capsules])
    db = np.stack([m.bias.grad.cpu().numpy() for m in layer.capsules])
    assert rel_err(dw, g['dweight']) < 1e-4
    assert rel_err(db, g['dbias']) < 1e-4


def test_dark_regroup_kernel(capsb):
    """caps_dark_regroup / _backward: bit-exact against the reference's own regroup (fixture) and, on
    random data and other shapes, against the oracle."""
    import os
    from conftest import GOLDEN_DIR, dark_pattern
    from oracle import routing_np as onp
    dev = torch.device('cuda')
    g = dict(np.load(os.path.join(GOLDEN_DIR, 'dark_regroup.npz')))
    B, Cch, grid = [int(v) for v in g['dims']]
    x = torch.from_numpy(dark_pattern((B, Cch, 4 * grid, 4 * grid), 7919, 8191)).to(dev).requires_grad_(True)
    u = capsb.dark_regroup(x, grid)
    assert tuple(u.shape) == (grid * grid * B, 2 * Cch, 8)
    assert np.array_equal(u.detach().cpu().numpy(), g['u'].astype(np.float32))
    u.backward(torch.from_numpy(dark_pattern(tuple(u.shape), 104729, 8179)).to(dev))
    assert np.array_equal(x.grad.cpu().numpy(), g['dx'].astype(np.float32))
    rng = np.random.default_rng(3)
    for (B2, C2, g2) in ((5, 64, 3), (1, 8, 1), (32, 256, 7)):
        xv = rng.standard_normal((B2, C2, 4 * g2, 4 * g2)).astype(np.float32)
        xt = torch.from_numpy(xv).to(dev).requires_grad_(True)
        ut = capsb.dark_regroup(xt, g2)
        assert np.array_equal(ut.detach().cpu().numpy(), onp.dark_regroup(xv, g2))
        duv = rng.standard_normal(tuple(ut.shape)).astype(np.float32)
        ut.backward(torch.from_numpy(duv).to(dev))
        assert np.array_equal(xt.grad.cpu().numpy().reshape(B2, C2, -1), onp.dark_regroup_bwd(duv, B2, C2, g2))
    with pytest.raises(RuntimeError):
        capsb.dark_regroup(torch.zeros(2, 12, 28, 28, device=dev), 7)        # Cch not a multiple of 8
    with pytest.raises(RuntimeError):
        capsb.dark_regroup(torch.zeros(2, 16, 28, 28), 7)                     # CPU tensor


def _rows_from_caps(caps):
    """[B,g,g,5] -> the routing layer's row order [g*g*B, 5] (row q*B + b), inverse of models.py:401"""
    B, g = caps.shape[0], caps.shape[1]
    return np.ascontiguousarray(caps.transpose(1, 2, 0, 3)).reshape(g * g * B, 5)


def test_dark_loss_kernel(capsb):
    """caps_dark_loss (value + gradient in one kernel) against the reference's darkcapsule_loss and autograd."""
    import os
    from conftest import GOLDEN_DIR
    from oracle import routing_np as onp
    dev = torch.device('cuda')
    g = dict(np.load(os.path.join(GOLDEN_DIR, 'dark_loss.npz')))
    B, grid = g['caps'].shape[0], g['caps'].shape[1]
    y = np.zeros((B, grid, grid, 48), dtype=np.float32)
    y[..., :5] = g['y5']
    v = torch.from_numpy(_rows_from_caps(g['caps'])).to(dev).view(grid * grid * B, 1, 1, 1, 5).requires_grad_(True)
    loss = capsb.dark_capsule_loss(v, torch.from_numpy(y).to(dev))
    assert abs(float(loss) - float(g['loss'])) < 2e-6 * abs(float(g['loss']))
    (3.0 * loss).backward()
    assert rel_err(v.grad.cpu().numpy().reshape(-1, 5), 3.0 * _rows_from_caps(g['dcaps'])) < 1e-5
    # a batch large enough for the two-level reduction, against the oracle
    rng = np.random.default_rng(8)
    B2, g2 = 300, 7
    caps = (0.5 * rng.standard_normal((B2, g2, g2, 5))).astype(np.float32)
    y2 = rng.uniform(0.05, 0.95, (B2, g2, g2, 6)).astype(np.float32)
    y2[..., 0] = (rng.uniform(size=(B2, g2, g2)) < 0.1)
    v2 = torch.from_numpy(_rows_from_caps(caps)).to(dev).requires_grad_(True)
    l2 = capsb.dark_capsule_loss(v2, torch.from_numpy(y2).to(dev))
    l2.backward()
    lo, go = onp.dark_loss(caps.astype(np.float64), y2.astype(np.float64))
    assert abs(float(l2) - lo) < 1e-5 * abs(lo)
    assert rel_err(v2.grad.cpu().numpy(), _rows_from_caps(go)) < 1e-5


def test_dark_capsule_chain(capsb):
    """regroup -> routing (512 -> 1 x 5, the DarkCapsuleNet head) -> darkcapsule loss, forward and backward,
    against the same chain assembled from the oracle's pieces."""
    from oracle import routing_np as onp
    dev = torch.device('cuda')
    rng = np.random.default_rng(12)
    B, grid, R = 3, 2, 3
    G = grid * grid
    x = (0.3 * rng.standard_normal((B, 256, 4 * grid, 4 * grid))).astype(np.float32)
    W = (0.1 * rng.standard_normal((512, 1, 8, 5))).astype(np.float32)
    y = rng.uniform(0.05, 0.95, (B, grid, grid, 48)).astype(np.float32)
    y[..., 0] = (rng.uniform(size=(B, grid, grid)) < 0.3)
    # oracle chain (fp64)
    u = onp.dark_regroup(x.astype(np.float64), grid)
    v, state = onp.routing_forward(u, W.astype(np.float64), R, return_state=True)        # [G*B,1,5]
    caps = v.reshape(grid, grid, B, 5).transpose(2, 0, 1, 3)                              # models.py:401
    loss, dcaps = onp.dark_loss(caps, y.astype(np.float64))
    gv = np.ascontiguousarray(dcaps.transpose(1, 2, 0, 3)).reshape(G * B, 1, 5)
    du, dW = onp.routing_backward(u, W.astype(np.float64), gv, R, state=state)
    dx = onp.dark_regroup_bwd(du, B, 256, grid).reshape(x.shape)
    # device chain
    layer = capsb.CapsuleLayer(None, n_caps=1, n_nodes=512, in_C=8, out_C=5, n_iter=R).to(dev)
    with torch.no_grad():
        layer.route_weights.copy_(torch.from_numpy(W)[None])
    xt = torch.from_numpy(x).to(dev).requires_grad_(True)
    out = layer(capsb.dark_regroup(xt, grid))                                             # [G*B,1,1,1,5]
    lt = capsb.dark_capsule_loss(out, torch.from_numpy(y).to(dev))
    lt.backward()
    assert abs(float(lt) - loss) < 1e-5 * max(abs(loss), 1.0)       # margin and coordinate terms cancel here: the sum is ~0.005
    assert rel_err(out.detach().cpu().numpy().reshape(G * B, 1, 5), v) < TOL_V
    assert rel_err(xt.grad.cpu().numpy(), dx) < TOL_G
    assert rel_err(layer.route_weights.grad[0].cpu().numpy(), dW) < TOL_G


def test_capsnet_cfg1_end_to_end(capsb):
    """BASELINE.json configs[0] shape: the reference's CapsuleNet (conv1 -> primary capsules -> routing -> scores,
    reference models.py:86-117) + capsule_loss + backward, rebuilt here from the drop-in layers in the reference's
    construction order under the same seed (the drop-in draws the same weights: tests/test_cabi_cpu.py), against
    what the unmodified reference produced on the CPU (tests/golden/capsnet_cfg1.npz)."""
    import os
    import torch.nn as nn
    import torch.nn.functional as F
    from conftest import GOLDEN_DIR
    dev = torch.device('cuda')
    g = dict(np.load(os.path.join(GOLDEN_DIR, 'capsnet_cfg1.npz')))
    B, xseed, st = [int(v) for v in g['dims']]
    torch.manual_seed(0)
    conv1 = nn.Conv2d(3, 256, 9)                                                     # models.py:90
    primary = capsb.CapsuleLayer(None, n_caps=8, n_nodes=-1, in_C=256, out_C=16, kernel=8, stride=2)
    digits = capsb.CapsuleLayer(None, n_caps=43, n_nodes=16 * 9 * 9, in_C=8, out_C=16)
    gen = torch.Generator().manual_seed(xseed)
    x = torch.rand(B, 3, 32, 32, generator=gen) * 2 - 1
    y = torch.randint(0, 43, (B,), generator=gen)
    assert abs(float(x.double().sum()) - float(g['x_checksum'])) < 1e-9 and np.array_equal(y.numpy(), g['y'])
    conv1, primary, digits = conv1.to(dev), primary.to(dev), digits.to(dev)
    torch.backends.cudnn.allow_tf32 = False
    xd = x.to(dev).requires_grad_(True)
    u = primary(F.relu(conv1(xd)))                                                   # models.py:113-114
    out, loss = digits.forward_margin_loss(u, y.to(dev))                             # models.py:115-117 + loss_fns.py:11-23
    scores = (out.squeeze() ** 2).sum(dim=-1) ** 0.5
    assert rel_err(scores.detach().cpu().numpy(), g['scores']) < 2e-5
    assert abs(float(loss) - float(g['loss'])) < 2e-5 * abs(float(g['loss']))
    loss.backward()
    tol = 2e-4
    assert rel_err(xd.grad.cpu().numpy(), g['dx']) < tol
    assert rel_err(conv1.weight.grad.reshape(-1)[::st].cpu().numpy(), g['conv1_w_probe']) < tol
    assert rel_err(conv1.bias.grad.cpu().numpy(), g['conv1_b']) < tol
    pw = torch.cat([m.weight.grad.reshape(-1) for m in primary.capsules])
    pb = torch.cat([m.bias.grad.reshape(-1) for m in primary.capsules])
    assert rel_err(pw[::st].cpu().numpy(), g['prim_w_probe']) < tol
    assert rel_err(pb.cpu().numpy(), g['prim_b']) < tol
    assert rel_err(digits.route_weights.grad.reshape(-1)[::st * 11].cpu().numpy(), g['route_w_probe']) < tol


@pytest.mark.gpu
def test_cuda_graph_step_matches_plain_calls(capsb):
    """GraphedStep (forward + margin loss + backward captured in one CUDA graph) replays to the same bits as the plain
    calls, for two different batches fed through its static buffers, and matches the fp64 oracle."""
    from oracle import routing_c as oc
    from oracle import routing_np as onp
    B, N, C, K, D, R = 64, 96, 43, 8, 16, 3
    dev = torch.device('cuda')
    u0, W, y0 = onp.make_inputs(B, N, C, K, D, seed=5)
    Wt = torch.from_numpy(W).to(dev)
    g = capsb.GraphedStep(B, N, C, K, D, R, Wt)
    for seed in (5, 6):
        u, _, y = onp.make_inputs(B, N, C, K, D, seed=seed)
        g.u.copy_(torch.from_numpy(u)); g.y.copy_(torch.from_numpy(y))
        g.replay()
        torch.cuda.synchronize()
        plain = cuda_step(capsb, u, W, y, R, want_c=False)
        assert np.array_equal(g.v.cpu().numpy(), plain['v'])
        assert np.array_equal(g.du.cpu().numpy(), plain['du'])
        assert np.array_equal(g.dW.cpu().numpy(), plain['dW'])
        assert float(g.loss) == plain['loss']
        ref = oc.routing_step(u.astype(np.float64), W.astype(np.float64), y, R, want_c=False)
        assert rel_err(g.v.cpu().numpy(), ref['v']) < 1e-5 and rel_err(g.dW.cpu().numpy(), ref['dW']) < 1e-4
