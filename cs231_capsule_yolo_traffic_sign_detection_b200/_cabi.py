"""ctypes binding of include/caps_routing.h (the C-ABI drop-in boundary).

This is the binding a maintainer of the reference would add behind `models.CapsuleLayer`
(INTEGRATION.md shows it stand-alone).  No torch types cross the boundary: only raw pointers,
ints and a stream handle.  There is no CPU fallback: if the library is missing or was not built,
importing a compute entry point raises."""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('CAPS_ROUTING_LIB') or os.path.join(HERE, 'libcaps_routing.so')   # env override: A/B experiments

ABI_VERSION = 3
MARGIN_SCRATCH_FLOATS = 2048

# every symbol include/caps_routing.h declares: name -> (restype, argtypes)
_vp, _i, _f, _sz, _l = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_size_t, ctypes.c_long
SYMBOLS = {
    'caps_abi_version': (_i, []),
    'caps_last_error': (ctypes.c_char_p, []),
    'caps_route_workspace_bytes': (_sz, [_i] * 7),
    'caps_route_forward': (_i, [_vp, _vp, _vp, _vp, _vp, _sz] + [_i] * 7 + [_vp]),
    'caps_route_backward': (_i, [_vp, _vp, _vp, _vp, _f, _vp, _vp, _vp, _vp, _sz] + [_i] * 6 + [_vp]),
    'caps_route_backward_ev': (_i, [_vp, _vp, _vp, _vp, _f, _vp, _vp, _vp, _vp, _sz] + [_i] * 6 + [_vp, _vp]),
    'caps_margin_loss': (_i, [_vp, _vp, _f, _vp, _vp, _vp, _i, _i, _i, _vp]),
    'caps_squash': (_i, [_vp, _vp, _l, _i, _vp]),
    'caps_squash_backward': (_i, [_vp, _vp, _vp, _l, _i, _vp]),
    'caps_primary_squash': (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    'caps_primary_squash_backward': (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    'caps_dark_regroup': (_i, [_vp, _vp, _i, _i, _i, _vp]),
    'caps_dark_regroup_backward': (_i, [_vp, _vp, _i, _i, _i, _vp]),
    'caps_dark_loss': (_i, [_vp, _vp, _f, _vp, _vp, _vp, _i, _i, _i, _vp]),
    'caps_route_step_host_scratch_bytes': (_sz, [_i] * 6),
    'caps_route_step_host': (_i, [_vp] * 8 + [_sz] + [_i] * 6 + [_vp]),
    'caps_host_pipe_scratch_bytes': (_sz, [_i] * 6),
    'caps_host_pipe_create': (_i, [_vp, _vp, _sz] + [_i] * 6),
    'caps_host_pipe_submit': (_i, [_vp, _vp, _vp]),
    'caps_host_pipe_step': (_i, [_vp] * 7),
    'caps_host_pipe_destroy': (_i, [_vp]),
    'caps_set_tuning': (_i, [ctypes.c_char_p, _i]),
    'caps_kernel_launch_count': (_l, []),
    'caps_profile_collect': (_i, [_vp, _vp, _i]),
    'caps_fma_peak': (_i, [_i, _vp, _vp, _vp]),
}

_lib = None


class CapsRoutingError(RuntimeError):
    pass


def lib():
    """Loads libcaps_routing.so (built in-tree by build.py).  Raises if it is not there:
    the product path never falls back to a CPU implementation."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise CapsRoutingError(
                'libcaps_routing.so is not built (%s). Run `python -c "import __graft_entry__ as g; '
                'g.build()"` or `python -m cs231_capsule_yolo_traffic_sign_detection_b200.build`. '
                'There is no CPU fallback.' % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)          # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        if L.caps_abi_version() != ABI_VERSION:
            raise CapsRoutingError('ABI mismatch: library %d, binding %d' % (L.caps_abi_version(), ABI_VERSION))
        _lib = L
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().caps_last_error().decode('utf-8', 'replace')
        raise CapsRoutingError('%s failed (code %d): %s' % (what, rc, msg))


def set_tuning(name, value):
    check(lib().caps_set_tuning(name.encode(), int(value)), 'caps_set_tuning')
