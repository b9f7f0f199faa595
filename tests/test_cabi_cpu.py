"""CPU-side checks of the boundary: the C-ABI library loads and exports every symbol the header
declares, argument validation works without a GPU, and the host-side module mirrors the
reference's interface (names, shapes, seeded init, error behaviour).  No compute calls."""
import ctypes
import os
import re
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def capsb():
    import __graft_entry__ as ge
    ge.build()                                      # nvcc cross-compiles without a GPU
    import cs231_capsule_yolo_traffic_sign_detection_b200 as m
    return m


def header_symbols():
    src = open(os.path.join(ROOT, 'include', 'caps_routing.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(caps_[a-z_0-9]+)\s*\(', src)))


def test_library_exports_every_declared_symbol(capsb):
    L = capsb._cabi.lib()
    names = header_symbols()
    assert 'caps_route_forward' in names and 'caps_route_backward' in names
    for n in names:
        assert hasattr(L, n), 'libcaps_routing.so does not export %s' % n
    assert sorted(capsb._cabi.SYMBOLS) == names, 'binding and header disagree'
    assert L.caps_abi_version() == capsb._cabi.ABI_VERSION


def test_workspace_sizing_and_unsupported_dims(capsb):
    L = capsb._cabi.lib()
    small = L.caps_route_workspace_bytes(64, 1152, 43, 8, 16, 3, 0)
    big = L.caps_route_workspace_bytes(64, 1152, 43, 8, 16, 3, 1)
    assert 0 < small < big
    # saved state is dominated by the (R-1) coupling arrays + (R-1) beta arrays + scratch
    assert big >= 5 * 64 * 1152 * 43 * 4
    assert L.caps_route_workspace_bytes(64, 1152, 43, 6, 16, 3, 1) == 0     # K != 8
    assert L.caps_route_workspace_bytes(64, 1152, 43, 8, 64, 3, 1) == 0     # D > 48
    assert L.caps_route_workspace_bytes(64, 1152, 43, 8, 16, 6, 1) == 0     # R > 5
    assert L.caps_route_workspace_bytes(64, 512, 1, 8, 5, 3, 1) > 0         # DarkCapsuleNet head
    assert L.caps_route_step_host_scratch_bytes(64, 1152, 43, 8, 16, 3) > big


def test_argument_errors_do_not_need_a_gpu(capsb):
    L = capsb._cabi.lib()
    rc = L.caps_route_forward(None, None, None, None, None, 0, 4, 8, 3, 6, 16, 3, 0, None)
    assert rc == -2 and b'not supported' in L.caps_last_error()
    rc = L.caps_route_forward(None, None, None, None, None, 0, 4, 8, 3, 8, 16, 3, 0, None)
    assert rc == -1 and b'null' in L.caps_last_error()
    buf = (ctypes.c_float * 64)()
    addr = ctypes.addressof(buf)
    rc = L.caps_route_forward(addr + 4, addr, addr, None, addr, 1 << 30, 4, 8, 3, 8, 16, 3, 0, None)
    assert rc == -1 and b'aligned' in L.caps_last_error()
    rc = L.caps_route_forward(addr, addr, addr, None, addr, 16, 4, 8, 3, 8, 16, 3, 0, None)
    assert rc == -3
    assert L.caps_set_tuning(b'nonsense', 1) == -1
    assert L.caps_set_tuning(b'spt', 3) == -1
    assert L.caps_set_tuning(b'spt', 0) == 0


def test_missing_library_fails_loudly(capsb, monkeypatch):
    monkeypatch.setattr(capsb._cabi, '_lib', None)
    monkeypatch.setattr(capsb._cabi, 'LIB_PATH', '/nonexistent/libcaps_routing.so')
    with pytest.raises(capsb._cabi.CapsRoutingError, match='no CPU fallback'):
        capsb._cabi.lib()


def test_cpu_tensors_raise_no_fallback(capsb):
    layer = capsb.CapsuleLayer(None, n_caps=3, n_nodes=4, in_C=8, out_C=16)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        layer(torch.zeros(2, 4, 8))


def test_module_mirrors_reference_interface(capsb):
    torch.manual_seed(7)
    layer = capsb.CapsuleLayer('params-object', n_caps=43, n_nodes=12, in_C=8, out_C=16)
    assert layer.params == 'params-object' and layer.n_iter == 3
    assert layer.n_nodes == 12 and layer.n_caps == 43
    assert list(layer.state_dict().keys()) == ['route_weights']
    assert tuple(layer.route_weights.shape) == (1, 12, 43, 8, 16)
    torch.manual_seed(7)                               # reference models.py:57-58: 0.1 * randn(1,N,C,K,D)
    assert torch.equal(layer.route_weights.detach(), 0.1 * torch.randn(1, 12, 43, 8, 16))
    # optional params.n_iter (the iteration-sweep knob); absent -> the constructor argument rules
    class WithIter:
        n_iter = 5
    assert capsb.CapsuleLayer(WithIter(), n_caps=3, n_nodes=4, in_C=8, out_C=16).n_iter == 5
    assert capsb.CapsuleLayer(object(), n_caps=3, n_nodes=4, in_C=8, out_C=16, n_iter=2).n_iter == 2
    prim = capsb.CapsuleLayer(None, n_caps=8, n_nodes=-1, in_C=16, out_C=4, kernel=3, stride=2)
    assert sorted(prim.state_dict().keys()) == sorted(
        ['capsules.%d.%s' % (i, w) for i in range(8) for w in ('weight', 'bias')])
    out = prim(torch.randn(2, 16, 9, 9))               # conv branch is stock PyTorch: runs on CPU
    assert tuple(out.shape) == (2, 4 * 4 * 4, 8)
    ref = torch.cat([c(torch.zeros(1, 16, 9, 9)).view(1, -1, 1) for c in prim.capsules], -1)
    assert ref.shape[1:] == out.shape[1:]


@pytest.mark.skipif(not os.path.exists('/root/reference/models.py'), reason='reference not mounted here')
def test_drop_in_into_reference_models(capsb):
    """Patch models.CapsuleLayer (class name resolved at construction) and check the
    reference's own CapsuleNet / DarkCapsuleNet build with identical parameter names, shapes and
    seeded values -- so reference checkpoints load (utils.py:52-60)."""
    sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
    import make_golden
    ref_models, _ = make_golden.import_reference()

    class P:
        device = 'cpu'; n_classes = 43; n_grid = 7
    orig = ref_models.CapsuleLayer
    try:
        for cls in (ref_models.CapsuleNet, ref_models.DarkCapsuleNet):
            torch.manual_seed(0)
            a = cls(P())
            ref_models.CapsuleLayer = capsb.CapsuleLayer
            torch.manual_seed(0)
            b = cls(P())
            ref_models.CapsuleLayer = orig
            sa, sb = a.state_dict(), b.state_dict()
            assert list(sa.keys()) == list(sb.keys())
            for k in sa:
                assert torch.equal(sa[k], sb[k]), k
            b.load_state_dict(sa, strict=True)
            assert isinstance(b.traffic_sign_capsules, capsb.CapsuleLayer)
    finally:
        ref_models.CapsuleLayer = orig
