"""Builds libcaps_routing.so IN-TREE with nvcc for sm_100a (cross-compiles without a GPU).

    python -m cs231_capsule_yolo_traffic_sign_detection_b200.build [--force] [--verbose]

The .so lands next to this file (git-ignored, but it travels to the GPU box with the gpurun
snapshot).  Translation units are compiled in parallel."""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, 'csrc')
INCLUDE = os.path.join(ROOT, 'include')
LIB = os.path.join(HERE, 'libcaps_routing.so')
OBJ_DIR = os.path.join(HERE, 'build')
UNITS = ['caps_api.cu', 'caps_pass.cu', 'caps_grad.cu', 'caps_pass_tc.cu', 'caps_grad_mma.cu', 'caps_sweep_fused.cu', 'caps_c1.cu']
HEADERS = [os.path.join(CSRC, 'caps_kernels.cuh'), os.path.join(CSRC, 'caps_internal.h'), os.path.join(CSRC, 'caps_tc_common.cuh'),
           os.path.join(INCLUDE, 'caps_routing.h')]
NVCC_FLAGS = ['-O3', '-std=c++17', '-lineinfo', '-gencode', 'arch=compute_100a,code=sm_100a',
              '-Xcompiler', '-fPIC', '-I', INCLUDE, '-I', CSRC]


def _nvcc():
    for cand in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    return 'nvcc'


def _digest():
    h = hashlib.sha256()
    for f in [os.path.join(CSRC, u) for u in UNITS] + HEADERS:
        h.update(open(f, 'rb').read())
    h.update(' '.join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    os.makedirs(OBJ_DIR, exist_ok=True)
    stamp = os.path.join(OBJ_DIR, 'stamp')
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == dig:
        return LIB
    nvcc = _nvcc()
    extra = ['-Xptxas', '-v'] if verbose else []

    def compile_unit(u):
        obj = os.path.join(OBJ_DIR, u.replace('.cu', '.o'))
        cmd = [nvcc] + NVCC_FLAGS + extra + ['-c', os.path.join(CSRC, u), '-o', obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose:
            open(os.path.join(OBJ_DIR, u + '.ptxas.log'), 'w').write(r.stderr)
        if r.returncode != 0:
            raise RuntimeError('nvcc failed for %s:\n%s' % (u, r.stderr[-4000:]))
        return obj

    with ThreadPoolExecutor(max_workers=len(UNITS)) as ex:
        objs = list(ex.map(compile_unit, UNITS))
    cmd = [nvcc, '-shared', '-gencode', 'arch=compute_100a,code=sm_100a', '-o', LIB] + objs
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('link failed:\n' + r.stderr[-4000:])
    open(stamp, 'w').write(dig)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='--verbose' in sys.argv))
