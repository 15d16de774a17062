"""One launch set of the fused linear train / eval kernels on a C5-like synthetic table, for ncu captures."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bear_b200 import _lib  # noqa: E402
from bear_b200._lib import lib, check, ptr  # noqa: E402

dev = torch.device('cuda', 0)
torch.cuda.set_device(dev)
n = int(os.environ.get('ROWS', 1 << 26))
lag = int(os.environ.get('LAG', 20))
regime = int(os.environ.get('REGIME', 0))
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
for _ in range(2):
    check(lib.bear_linear_train_step(ptr(kmers), ptr(counts), stride, 0, n, lag, ptr(mat), ptr(hs), 1.0, 0, ptr(flat), None,
                                     ptr(ws), _lib.stream()))
    check(lib.bear_eval_step(ptr(kmers), ptr(counts), None, stride, 0, n, lag, _lib.HEAD_LINEAR, ptr(mat), ptr(h), 1,
                             ptr(alpha), 3, 7, 0, ptr(eacc), ptr(ws), _lib.stream()))
torch.cuda.synchronize()
print('ok', float(flat[0]))
if os.environ.get('BMM'):
    acc3 = torch.zeros(3, dtype=torch.float64, device=dev)
    for _ in range(2):
        check(lib.bear_bmm_likelihood(ptr(counts), stride, 0, n, 1, 5, ptr(alpha), 3, ptr(acc3), ptr(ws), _lib.stream()))
    torch.cuda.synchronize()
