"""Builds libbear_b200.so (C-ABI, sm_100a only) in-tree with nvcc.

    python -m bear_b200.build [--force] [--verbose]

The shared library lands next to this file so that it travels with the source tree; nothing is
JIT-compiled at run time and there is no fallback when it is missing.  Sources are compiled in parallel
into bear_b200/_build/ (one object per translation unit, recompiled only when that unit or a header
changed).  The digest of sources + flags is compiled INTO the library (bear_build_digest()), so a stale
binary is detected from the binary itself, never from a side file.
"""
import argparse
import concurrent.futures
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libbear_b200.so')
OBJ = os.path.join(HERE, '_build')

SOURCES = ['bear_pack.cpp', 'bear_dense.cu', 'bear_fused.cu', 'bear_eval_misc.cu', 'bear_eval_lin.cu', 'bear_eval_ref.cu',
           'bear_train.cu', 'bear_heads.cu', 'bear_count.cu', 'bear_cnn.cu']
HEADERS = ['bear_common.cuh', 'bear_dm_row.cuh', 'bear_host.h', 'bear_linear_head.cuh', 'bear_sm100.cuh', 'bear_eval.cuh', 'bear_rank.h']

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a',
    '-O3', '-lineinfo', '--std=c++17',
    '-Xcompiler', '-fPIC',
    '-Xptxas', '-v',
]

# extra nvcc flags for kernel experiments, e.g. BEAR_NVCC_EXTRA=-DBEAR_TRAIN_EXPERIMENTS (part of the build digest)
NVCC_FLAGS += os.environ.get('BEAR_NVCC_EXTRA', '').split()


def _nvcc():
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError('nvcc not found; libbear_b200.so cannot be built')


def _read(path):
    with open(path, 'rb') as fh:
        return fh.read()


def _common_digest(only=None):
    """Hash of the headers (all of them, or those in `only`), the public header and the compiler flags."""
    h = hashlib.sha256()
    for f in HEADERS:
        if only is None or f in only:
            h.update(_read(os.path.join(CSRC, f)))
    h.update(_read(os.path.join(ROOT, 'include', 'bear_b200.h')))
    h.update(' '.join(NVCC_FLAGS).encode())
    return h


def _includes(name, seen=None):
    """Project headers a source file includes, transitively."""
    seen = set() if seen is None else seen
    for line in _read(os.path.join(CSRC, name)).decode().splitlines():
        line = line.strip()
        if line.startswith('#include "'):
            inc = line.split('"')[1]
            if inc in HEADERS and inc not in seen:
                seen.add(inc)
                _includes(inc, seen)
    return seen


def _digest():
    """Digest of everything the binary is made from: sources, headers, the public header, compiler flags."""
    h = _common_digest()
    for f in SOURCES:
        h.update(_read(os.path.join(CSRC, f)))
    return h.hexdigest()


def library_digest(path=LIB):
    """The digest compiled into an existing library (read from the file, not through dlopen: a handle of an older
    build may already be loaded in this process), or None (missing library / pre-digest build)."""
    if not os.path.exists(path):
        return None
    data = _read(path)
    i = data.find(b'BEAR_BUILD_DIGEST:')
    if i < 0:
        return None
    return data[i + 18:i + 18 + 64].decode('ascii', 'replace')


def build(force=False, verbose=False):
    digest = _digest()
    if not force and library_digest() == digest:
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()
    inc = ['-I', os.path.join(ROOT, 'include'), '-I', CSRC]
    logs = {}

    def compile_one(src):
        h = _common_digest(_includes(src))
        h.update(_read(os.path.join(CSRC, src)))
        if src == 'bear_pack.cpp':
            h.update(digest.encode())                       # carries bear_build_digest()
        tag = h.hexdigest()
        obj = os.path.join(OBJ, src + '.o')
        stamp = obj + '.digest'
        if not force and os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == tag:
            return obj, None
        cmd = [nvcc] + NVCC_FLAGS + inc + ['-DBEAR_BUILD_DIGEST="%s"' % digest, '-c', os.path.join(CSRC, src), '-o', obj]
        proc = subprocess.run(cmd, capture_output=True, text=True)
        logs[src] = proc.stdout + proc.stderr
        if proc.returncode != 0:
            return obj, logs[src]
        with open(obj + '.log', 'w') as fh:                  # ptxas -v output of this unit (registers, spills)
            fh.write(logs[src])
        with open(stamp, 'w') as fh:
            fh.write(tag)
        return obj, None

    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as pool:
        results = list(pool.map(compile_one, SOURCES))
    if verbose:
        for src in SOURCES:
            sys.stderr.write(logs.get(src, ''))
    errors = [e for _, e in results if e]
    if errors:
        if not verbose:
            sys.stderr.write('\n'.join(errors))
        raise RuntimeError('nvcc failed building libbear_b200.so')
    cmd = [nvcc, '-gencode', 'arch=compute_100a,code=sm_100a', '-shared', '-o', LIB] + [o for o, _ in results]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
        raise RuntimeError('nvcc failed linking libbear_b200.so')
    with open(os.path.join(OBJ, 'ptxas.log'), 'w') as fh:      # registers / spills per kernel, for tools/ and profiles/
        for src in SOURCES:
            if os.path.exists(os.path.join(OBJ, src + '.o.log')):
                fh.write('## %s\n%s' % (src, open(os.path.join(OBJ, src + '.o.log')).read()))
    if library_digest() != digest:
        raise RuntimeError('libbear_b200.so was built but does not report the expected source digest')
    return LIB


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--force', action='store_true')
    ap.add_argument('--verbose', action='store_true')
    a = ap.parse_args()
    print(build(a.force, a.verbose))
