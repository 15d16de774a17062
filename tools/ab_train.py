"""Quick A/B timing of the fused linear train / eval kernels (CUDA events) for kernel experiments."""
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
out = []
for lag, regime, name in ((20, 0, 'lag20'), (20, 2, 'lag20-sorted'), (13, 0, 'lag13'), (20, 1, 'lag20-dense')):
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

    def train():
        check(lib.bear_linear_train_step(ptr(kmers), ptr(counts), stride, 0, n, lag, ptr(mat), ptr(hs), 1.0, 0, ptr(flat), None,
                                         ptr(ws), _lib.stream()))

    def evalk():
        check(lib.bear_eval_step(ptr(kmers), ptr(counts), None, stride, 0, n, lag, _lib.HEAD_LINEAR, ptr(mat), ptr(h), 1,
                                 ptr(alpha), 3, 7, ptr(eacc), ptr(ws), _lib.stream()))
    for fn, kn in ((train, 'train'), (evalk, 'eval')):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        cyc = None
        if kn == 'train' and hasattr(lib, 'bear_debug_train_cycles'):      # built with -DBEAR_TRAIN_EXPERIMENTS
            import ctypes
            cyc = (ctypes.c_ulonglong * 5)()
            lib.bear_debug_train_cycles(cyc)                                # clear
            fn()
            lib.bear_debug_train_cycles(cyc)
            ctas = max(1, (cyc[3] + cyc[4]) // 16)
            out.append('%s busy: producers %.2f, consumers %.2f of the kernel loop' % (
                name, cyc[0] / max(1, cyc[2] * cyc[3] / ctas), cyc[1] / max(1, cyc[2] * cyc[4] / ctas)))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        out.append('%s %s: %.3f ms (%.3e rows/s)' % (name, kn, ms, n / ms * 1e3))
    del kmers, counts
print(tag, ' | '.join(out))
