"""Builds ``liblcr_b200.so`` in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
REPO = os.path.dirname(HERE)
OBJ_DIR = os.path.join(REPO, 'build', 'obj')
LIB_PATH = os.path.join(HERE, 'liblcr_b200.so')

NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC', '-I', os.path.join(REPO, 'include')]


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith('.cu'))


def _digest(paths):
    h = hashlib.sha256()
    h.update(' '.join(NVCC_FLAGS).encode())
    for p in paths:
        with open(p, 'rb') as f:
            h.update(f.read())
    return h.hexdigest()


def build_library(force=False, verbose=False):
    os.makedirs(OBJ_DIR, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(('.cuh', '.h'))]
    headers.append(os.path.join(REPO, 'include', 'lcr_b200.h'))
    objs, rebuilt = [], False
    procs = []
    for src in _sources():
        sp = os.path.join(CSRC, src)
        obj = os.path.join(OBJ_DIR, src[:-3] + '.o')
        stamp = obj + '.sha'
        dig = _digest([sp] + headers)
        objs.append(obj)
        if not force and os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == dig:
            continue
        cmd = [NVCC] + NVCC_FLAGS + ['-c', sp, '-o', obj]
        if verbose:
            print(' '.join(cmd), file=sys.stderr)
        procs.append((subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT), stamp, dig, src))
        rebuilt = True
    for p, stamp, dig, src in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError('nvcc failed for %s:\n%s' % (src, out.decode()))
        with open(stamp, 'w') as f:
            f.write(dig)
    if rebuilt or not os.path.exists(LIB_PATH):
        cmd = [NVCC, '-shared', '-gencode', 'arch=compute_100a,code=sm_100a', '-o', LIB_PATH] + objs
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
        if r.returncode != 0:
            raise RuntimeError('link failed:\n%s' % r.stdout.decode())
    return LIB_PATH


if __name__ == '__main__':
    print(build_library(force='--force' in sys.argv, verbose=True))
