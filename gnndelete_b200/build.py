"""In-tree nvcc build of ``libgnndelete_b200.so`` for sm_100a (cross-compiles without a GPU).

    python -m gnndelete_b200.build [--force] [--verbose]
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OBJ_DIR = os.path.join(CSRC, '_build')
LIB = os.path.join(HERE, 'libgnndelete_b200.so')
HEADER = os.path.join(os.path.dirname(HERE), 'include', 'gnndelete_b200.h')

NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
ARCH = ['-gencode', 'arch=compute_100a,code=sm_100a']
CFLAGS = ['-O3', '-std=c++17', '-lineinfo', '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr']


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cu'))


def _deps_mtime():
    m = os.path.getmtime(HEADER)
    for f in os.listdir(CSRC):
        if f.endswith(('.cuh', '.h')):
            m = max(m, os.path.getmtime(os.path.join(CSRC, f)))
    return m


def _run(cmd, verbose):
    if verbose:
        print(' '.join(cmd), flush=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f'build failed: {" ".join(cmd)}\n{r.stdout}\n{r.stderr}')
    if verbose and (r.stdout or r.stderr):
        print(r.stdout, r.stderr)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ_DIR, exist_ok=True)
    hdr = _deps_mtime()
    jobs, objs = [], []
    for src in _sources():
        obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + '.o')
        objs.append(obj)
        stale = force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr)
        if stale:
            jobs.append([NVCC, *ARCH, *CFLAGS, '-c', src, '-o', obj])
    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(lambda c: _run(c, verbose), jobs))
    if jobs or not os.path.exists(LIB):
        _run([NVCC, *ARCH, '-shared', '-cudart', 'static', '-o', LIB, *objs], verbose)
    return LIB


if __name__ == '__main__':
    path = build(force='--force' in sys.argv, verbose='--verbose' in sys.argv)
    print(path)
