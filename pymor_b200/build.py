"""Build libpymor_b200.so (sm_100a only) in-tree with nvcc.

The shared library is the C-ABI drop-in boundary declared in include/pymor_b200.h.  It is built
next to the sources (pymor_b200/lib/) so that it travels with the repository snapshot to the GPU box;
nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / 'csrc'
LIB_DIR = PKG_DIR / 'lib'
LIB_PATH = LIB_DIR / 'libpymor_b200.so'
SOURCES = ['ctx.cu', 'blas1.cu', 'mgs.cu', 'gram.cu', 'gram2.cu', 'lincomb.cu', 'csr.cu', 'comm.cu', 'peak.cu']

NVCC_FLAGS = [
    '-O3', '-std=c++17', '-lineinfo',
    '-gencode', 'arch=compute_100a,code=sm_100a',
    '-Xcompiler', '-fPIC',
    '--expt-relaxed-constexpr',
]


def _nvcc() -> str:
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError('nvcc not found; cannot build libpymor_b200.so')


def _digest() -> str:
    h = hashlib.sha256()
    for f in sorted(CSRC.glob('*.cu*')) + [PKG_DIR.parent / 'include' / 'pymor_b200.h']:
        h.update(f.name.encode())
        h.update(f.read_bytes())
    h.update(' '.join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every CUDA source for sm_100a and link the shared library.  Returns its path."""
    LIB_DIR.mkdir(exist_ok=True)
    stamp = LIB_DIR / 'build.sha256'
    digest = _digest()
    if not force and LIB_PATH.exists() and stamp.exists() and stamp.read_text().strip() == digest:
        return LIB_PATH
    nvcc = _nvcc()
    objs = []
    procs = []
    for src in SOURCES:
        obj = LIB_DIR / (src.replace('.cu', '.o'))
        cmd = [nvcc, *NVCC_FLAGS, '-c', str(CSRC / src), '-o', str(obj)]
        if verbose:
            cmd.insert(1, '-Xptxas')
            cmd.insert(2, '-v')
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(str(obj))
    failed = False
    for src, pr in procs:
        out, _ = pr.communicate()
        if pr.returncode != 0 or verbose:
            sys.stderr.write(f'--- nvcc {src} ---\n{out}\n')
        failed |= pr.returncode != 0
    if failed:
        raise RuntimeError('nvcc failed, see output above')
    cmd = [nvcc, '-shared', '-o', str(LIB_PATH), *objs, '-gencode', 'arch=compute_100a,code=sm_100a', '-ldl']
    subprocess.run(cmd, check=True)
    stamp.write_text(digest)
    return LIB_PATH


if __name__ == '__main__':
    p = build(force='--force' in sys.argv, verbose='-v' in sys.argv)
    print(p)
