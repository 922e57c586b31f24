"""Build libdhr_b200.so (hand-written CUDA for sm_100a + C ABI) in-tree with nvcc.

    python -m dhr_b200.build            # incremental
    python -m dhr_b200.build --force

The library links only against cudart (no torch); it is loaded through ctypes by
dhr_b200._cabi.  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import argparse
import concurrent.futures
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OBJ = os.path.join(HERE, 'build')
LIB = os.path.join(HERE, 'lib', 'libdhr_b200.so')
SOURCES = ['index.cu', 'search.cu', 'scan_dispatch.cu', 'topk.cu', 'dense_tile.cu', 'lex_tile.cu', 'densify.cu', 'trec.cu']
GROUPS = [1, 2, 3, 4, 5, 6, 7, 8]          # scan_inst.cu is compiled once per G (values per slice)
ARCH = ['-gencode', 'arch=compute_100a,code=sm_100a']


def nvcc():
    exe = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(exe):
        raise RuntimeError('nvcc not found')
    return exe


def _newest_header():
    ts = [os.path.getmtime(os.path.join(CSRC, f)) for f in os.listdir(CSRC) if f.endswith(('.h', '.cuh'))]
    ts.append(os.path.getmtime(os.path.join(HERE, '..', 'include', 'dhr_b200.h')))
    return max(ts)


def _compile(job, force, verbose):
    src, defs, tag = job
    obj = os.path.join(OBJ, os.path.splitext(src)[0] + tag + '.o')
    srcp = os.path.join(CSRC, src)
    if not force and os.path.exists(obj) and os.path.getmtime(obj) >= max(os.path.getmtime(srcp), _newest_header()):
        return obj, ''
    cmd = [nvcc(), *ARCH, '-O3', '-std=c++17', '-lineinfo', '-Xcompiler', '-fPIC', '-Xptxas', '-v' if verbose else '-w',
           '--expt-relaxed-constexpr', *defs, '-c', srcp, '-o', obj]
    p = subprocess.run(cmd, capture_output=True, text=True)
    if p.returncode != 0:
        raise RuntimeError('nvcc failed for %s:\n%s\n%s' % (src, p.stdout, p.stderr))
    return obj, p.stderr


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    jobs = [('scan_inst.cu', ['-DDHR_G=%d' % g], '_g%d' % g) for g in GROUPS] + [(s, [], '') for s in SOURCES]
    with concurrent.futures.ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        results = list(ex.map(lambda j: _compile(j, force, verbose), jobs))
    objs = [r[0] for r in results]
    if verbose:
        for r in results:
            sys.stderr.write(r[1])
    if force or not os.path.exists(LIB) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs):
        cmd = [nvcc(), *ARCH, '-shared', '--cudart', 'static', '-o', LIB, *objs]
        p = subprocess.run(cmd, capture_output=True, text=True)
        if p.returncode != 0:
            raise RuntimeError('link failed:\n%s\n%s' % (p.stdout, p.stderr))
    return LIB


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--force', action='store_true')
    ap.add_argument('--verbose', action='store_true')
    a = ap.parse_args()
    print(build(a.force, a.verbose))
