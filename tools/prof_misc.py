"""One launch set of the pure streaming kernels (bmm, decode_onehot, unpack_counts) on a C5-like synthetic table, for ncu."""
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
lag, G = 20, int(os.environ.get('G', 1))
stride = (n + 3) // 4 * 4
kmers = torch.empty(stride, dtype=torch.int64, device=dev)
counts = torch.empty((G, 5, stride), dtype=torch.int32, device=dev)
check(lib.bear_synth_table(ptr(kmers), ptr(counts), stride, 0, n, lag, G, 20, 0, 10, _lib.stream()))
ws = torch.empty(lib.bear_workspace_doubles(n, lag, 0), dtype=torch.float64, device=dev)
alpha = torch.tensor([0.1, 1.0, 10.0], dtype=torch.float64, device=dev)
acc3 = torch.zeros(3 * G, dtype=torch.float64, device=dev)
m = min(n, 1 << 24)
onehot = torch.empty((m, lag, 5), dtype=torch.float64, device=dev)
dense = torch.empty((m, G, 5), dtype=torch.float64, device=dev)
for _ in range(2):
    check(lib.bear_bmm_likelihood(ptr(counts), stride, 0, n, G, 5, ptr(alpha), 3, ptr(acc3), ptr(ws), _lib.stream()))
    check(lib.bear_decode_onehot(ptr(kmers), m, lag, 0, ptr(onehot), _lib.stream()))
    check(lib.bear_unpack_counts(ptr(counts), stride, 0, m, G, 5, ptr(dense), _lib.stream()))
torch.cuda.synchronize()
print('ok')
