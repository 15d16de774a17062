"""Builds libbear_b200.so (C-ABI, sm_100a only) in-tree with nvcc.

    python -m bear_b200.build [--force] [--verbose]

The shared library lands next to this file so that it travels with the source tree; nothing is
JIT-compiled at run time and there is no fallback when it is missing.
"""
import argparse
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libbear_b200.so')
STAMP = os.path.join(HERE, '.libbear_b200.stamp')

SOURCES = ['bear_pack.cpp', 'bear_dense.cu', 'bear_fused.cu', 'bear_train.cu', 'bear_heads.cu', 'bear_count.cu', 'bear_cnn.cu']
HEADERS = ['bear_common.cuh', 'bear_dm_row.cuh', 'bear_host.h', 'bear_linear_head.cuh', 'bear_sm100.cuh']

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a',
    '-O3', '-lineinfo', '--std=c++17',
    '-Xcompiler', '-fPIC', '-shared',
    '-Xptxas', '-v',
]


# extra nvcc flags for kernel experiments, e.g. BEAR_NVCC_EXTRA=-DBEAR_TRAIN_EXPERIMENTS (part of the build digest)
NVCC_FLAGS += os.environ.get('BEAR_NVCC_EXTRA', '').split()


def _nvcc():
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError('nvcc not found; libbear_b200.so cannot be built')


def _digest():
    h = hashlib.sha256()
    files = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.join(ROOT, 'include', 'bear_b200.h')]
    for f in files:
        with open(f, 'rb') as fh:
            h.update(fh.read())
    h.update(' '.join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    digest = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP):
        with open(STAMP) as fh:
            if fh.read().strip() == digest:
                return LIB
    cmd = [_nvcc()] + NVCC_FLAGS + ['-I', os.path.join(ROOT, 'include'), '-I', CSRC,
                                    '-o', LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
    if proc.returncode != 0:
        raise RuntimeError('nvcc failed building libbear_b200.so')
    with open(os.path.join(HERE, 'build_ptxas.log'), 'w') as fh:
        fh.write(proc.stdout + proc.stderr)
    with open(STAMP, 'w') as fh:
        fh.write(digest)
    return LIB


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--force', action='store_true')
    ap.add_argument('--verbose', action='store_true')
    a = ap.parse_args()
    print(build(a.force, a.verbose))
