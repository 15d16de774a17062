"""A/B timing of the fused linear train / eval kernels on synthetic tables (CUDA events, inputs >> L2).
If the library also exports bear_linear_train_step_legacy (a previous kernel kept for comparison) it is timed too and
its outputs are compared."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bear_b200 import _lib  # noqa: E402
from bear_b200._lib import lib, check, ptr  # noqa: E402

dev = torch.device('cuda', 0)
torch.cuda.set_device(dev)
n = int(os.environ.get('ROWS', 1 << 27))
tag = os.environ.get('TAG', '')
cases = os.environ.get('CASES', 'lag20,lag20-sorted,lag13,lag20-dense').split(',')
legacy = getattr(ctypes.CDLL(_lib.LIB_PATH), 'bear_linear_train_step_legacy', None)
if legacy is not None:
    legacy.restype = ctypes.c_int
    legacy.argtypes = lib.bear_linear_train_step.argtypes
out = []
for lag, regime, name in ((20, 0, 'lag20'), (20, 2, 'lag20-sorted'), (13, 0, 'lag13'), (20, 1, 'lag20-dense')):
    if name not in cases:
        continue
    stride = (n + 3) // 4 * 4
    kmers = torch.empty(stride, dtype=torch.int64, device=dev)
    counts = torch.empty((1, 5, stride), dtype=torch.int32, device=dev)
    check(lib.bear_synth_table(ptr(kmers), ptr(counts), stride, 0, n, lag, 1, 20, regime, 10, _lib.stream()))
    ws = torch.empty(lib.bear_workspace_doubles(n, lag, 0), dtype=torch.float64, device=dev)
    mat = (torch.randn(lag, 5, 5, dtype=torch.float64, device=dev) * 0.05).contiguous()
    hs = torch.zeros(1, dtype=torch.float64, device=dev)
    flat = torch.zeros(2 + lag * 25, dtype=torch.float64, device=dev)
    h = torch.ones(1, dtype=torch.float64, device=dev)
    alpha = torch.tensor([0.1, 1.0, 10.0], dtype=torch.float64, device=dev)
    eacc = torch.zeros(11, dtype=torch.float64, device=dev)

    def train(fn=lib.bear_linear_train_step):
        check(fn(ptr(kmers), ptr(counts), stride, 0, n, lag, ptr(mat), ptr(hs), 1.0, 0, ptr(flat), None, ptr(ws), _lib.stream()))

    def evalk():
        check(lib.bear_eval_step(ptr(kmers), ptr(counts), None, stride, 0, n, lag, _lib.HEAD_LINEAR, ptr(mat), ptr(h), 1,
                                 ptr(alpha), 3, 7, 0, ptr(eacc), ptr(ws), _lib.stream()))
    fns = [(train, 'train'), (evalk, 'eval')]
    if legacy is not None:
        flat.zero_()
        train()
        new = flat.clone()
        flat.zero_()
        train(legacy)
        old = flat.clone()
        err = float((new[2:] - old[2:]).abs().max() / old[2:].abs().max())
        out.append('%s: new vs legacy gradient %.2e (rel. to largest), loss %.2e, dh %.2e' % (
            name, err, abs(float(new[0] - old[0]) / float(old[0])), abs(float(new[1] - old[1]) / float(old[1]))))
        fns.insert(1, (lambda: train(legacy), 'train-legacy'))
    for fn, kn in fns:
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        dbg = getattr(ctypes.CDLL(_lib.LIB_PATH), 'bear_debug_t3', None)       # built with -DBEAR_T3_DEBUG
        if kn == 'train' and dbg is not None:
            buf = (ctypes.c_ulonglong * 8)()
            dbg(buf)
            fn()
            dbg(buf)
            out.append('%s issuer: %.0f cycles per product in the loop, %.0f of them issuing; %.2f sweeps per product (%.2f empty)' % (
                name, buf[0] / max(buf[4], 1), buf[1] / max(buf[4], 1), buf[2] / max(buf[4], 1), buf[3] / max(buf[4], 1)))
            if buf[5]:       # (pipeline wait counters: only in the tools/experiments/train_mbarrier_issuer.patch build)
                out.append('%s producer, cycles per tile: %.0f waiting for the input stage, %.0f waiting for the slab, %.0f writing operands, %.0f in the proxy fence' % (
                    name, buf[3] / max(buf[4], 1), buf[5] / max(buf[4], 1), buf[7] / max(buf[4], 1), buf[6] / max(buf[4], 1)))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        out.append('%s %s: %.3f ms (%.3e rows/s)' % (name, kn, ms, n / ms * 1e3))
    del kmers, counts
print(tag, '\n'.join(out))
