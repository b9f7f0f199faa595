"""ctypes wrapper around the plain-C oracle (oracle/routing_c.c).  TEST INFRASTRUCTURE ONLY.

`build()` compiles it with gcc into oracle/_ref/libcaps_oracle.so (git-ignored; travels to the
GPU box with the gpurun snapshot).  Used by tests at sizes where the torch oracle (65 MB of
autograd state per sample) is too heavy, and as an fp64 tie-breaker."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, '_ref', 'libcaps_oracle.so')
_lib = None


def build(force=False):
    src = os.path.join(HERE, 'routing_c.c')
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.check_call(['make', '-s', '-C', HERE, 'all'])
    return LIB


def _load():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(LIB)
    return _lib


CHUNK = 8


def routing_step(u, W, y=None, n_iter=3, grad_v_extra=None, inv_batch=None, backward=True,
                 want_c=True, threads=None):
    """u [B,N,K], W [N,C,K,D] (float32 or float64, same dtype), y [B] int64 or None.
    Returns dict(v, c, loss, du, dW) like oracle.routing_np.routing_step.
    The batch is cut into fixed 8-sample chunks run on `threads` host threads (ctypes drops the
    GIL); per-chunk dW partials are added in chunk order -> thread-count independent bits."""
    from concurrent.futures import ThreadPoolExecutor
    dt = u.dtype
    assert dt in (np.float32, np.float64) and W.dtype == dt
    fn = getattr(_load(), 'caps_oracle_step_f32' if dt == np.float32 else 'caps_oracle_step_f64')
    ct = ctypes.c_float if dt == np.float32 else ctypes.c_double
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.c_void_p] * 4 + [ct] + [ctypes.c_int] * 6 + [ctypes.c_void_p] * 5 + [ctypes.c_int]
    B, N, K = u.shape
    _, C, _, D = W.shape
    u = np.ascontiguousarray(u); W = np.ascontiguousarray(W)
    v = np.empty((B, C, D), dt)
    c = np.empty((B, N, C), dt) if want_c else None
    du = np.empty((B, N, K), dt) if backward else None
    dW = np.zeros((N, C, K, D), dt) if backward else None
    if inv_batch is None:
        inv_batch = 1.0 / max(B, 1)
    if y is not None:
        y = np.ascontiguousarray(y, dtype=np.int64)
    if grad_v_extra is not None:
        grad_v_extra = np.ascontiguousarray(grad_v_extra, dtype=dt)
    threads = threads or os.cpu_count() or 1

    def p(a, lo=0):
        return None if a is None else a[lo:].ctypes.data_as(ctypes.c_void_p)

    def run(lo):
        hi = min(B, lo + CHUNK)
        part = np.zeros_like(dW) if backward else None
        loss = np.zeros((1,), dt)
        rc = fn(p(u, lo), p(W), p(y, lo), p(grad_v_extra, lo), ct(inv_batch), hi - lo, N, C, K, D,
                n_iter, p(v, lo), p(c, lo), p(loss), p(du, lo), p(part), int(backward))
        if rc != 0:
            raise ValueError('caps_oracle_step: bad arguments')
        return loss[0], part

    total = dt.type(0)
    with ThreadPoolExecutor(max_workers=threads) as ex:
        for l, part in ex.map(run, range(0, B, CHUNK)):
            total = total + l
            if backward:
                dW += part
    return dict(v=v, c=c, loss=total, du=du, dW=dW)
