"""Per-kernel throughput of the count-streaming kernels on synthetic tables (CUDA events, inputs >> L2).
Prints one JSON object per kernel: rows/s, algorithmic GB/s and fraction of the measured HBM peak."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bear_b200 import _lib  # noqa: E402
from bear_b200._lib import lib, check, ptr  # noqa: E402


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


def main():
    dev = torch.device('cuda', 0)
    torch.cuda.set_device(dev)
    peak = 6650.0
    try:
        peak = float(json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'])
    except Exception:
        pass
    K = int(os.environ.get('ROWS', 1 << 28))
    out = []

    def report(name, secs, rows, bytes_per_row, note=''):
        gbs = rows * bytes_per_row / secs / 1e9
        out.append({'kernel': name, 'rows': rows, 'ms': secs * 1e3, 'rows_per_s': rows / secs, 'bytes_per_row': bytes_per_row,
                    'achieved_gbs': gbs, 'frac_of_measured_hbm_peak': gbs / peak, 'note': note})
        print(json.dumps(out[-1]), flush=True)

    only = os.environ.get('ONLY', '')

    def cnn_section():
        """C4: CNN head (filter_width 3, 30 filters, layer width 16), lag 13, 4 groups -- fused kernels vs the
        torch-op explicit route (the plugin path) on the same rows."""
        from bear_b200 import ar_funcs, _engine as eng, dataloader as dl
        lag, G, W, F, H1 = 13, 4, 3, 30, 16
        n = int(os.environ.get('CNN_ROWS', 1 << 22))
        stride = (n + 3) // 4 * 4
        kmers = torch.empty(stride, dtype=torch.int64, device=dev)
        counts = torch.empty((G, 5, stride), dtype=torch.int32, device=dev)
        check(lib.bear_synth_table(ptr(kmers), ptr(counts), stride, 0, n, lag, G, 24, 1, 10, _lib.stream()))
        table = dl.KmerTable.from_device(kmers, counts, n, lag, 'dna')
        torch.manual_seed(0)
        af, params = ar_funcs.make_ar_func_cnn(lag, 4, filter_width=W)
        block = torch.cat([p.reshape(-1) for p in params]).contiguous()
        npar = block.numel()
        ws = torch.empty(lib.bear_workspace_doubles(n, lag, npar), dtype=torch.float64, device=dev)
        hs = torch.zeros(1, dtype=torch.float64, device=dev)
        flat = torch.zeros(2 + npar, dtype=torch.float64, device=dev)
        flops = 3 * 2 * (lag - W + 1) * F * H1        # three dense-layer-1 contractions per row
        for ar in (0, 1):
            t = timed(lambda: check(lib.bear_cnn_train_step(ptr(kmers), ptr(counts), stride, 0, n, lag, W, F, H1, ptr(block), ptr(hs),
                                                            1.0, ar, ptr(flat), None, ptr(ws), _lib.stream())), reps=3, warm=1)
            report('cnn_kernel<TRAIN_%s> [C4 dense lag13 G4 W3 F30 H16]' % ('AR' if ar else 'BEAR'), t, n, 28,
                   'compute-bound: %.2f TFLOP/s of FP64 tensor-core work (dense layer 1 fwd + 2 bwd)' % (flops * n / t / 1e12))
        f = torch.empty((n, 5), dtype=torch.float64, device=dev)
        t = timed(lambda: check(lib.bear_cnn_head_forward(ptr(kmers), 0, n, lag, W, F, H1, ptr(block), ptr(f), _lib.stream())), reps=3, warm=1)
        report('cnn_kernel<FWD> [C4 dense lag13 G4 W3 F30 H16]', t, n, 48, 'reads 8 B k-mer, writes 40 B f')
        m = min(n, 1 << 20)
        fp = eng.FlatParams([hs.reshape(())] + [p.requires_grad_(True) for p in params])

        def torch_route():
            eng.explicit_train_step(table, 0, 0, m, 1.0, False, fp, ws, lambda c0, cn: eng.explicit_f(af, table, c0, cn))
        t = timed(torch_route, reps=2, warm=1)
        report('torch-op explicit route, CNN train step [C4 dense lag13 G4 W3 F30 H16]', t, m, 28,
               'plugin path: decode_onehot -> torch ops -> bear_dm_train_step_explicit -> autograd')

    if only in ('', 'cnn'):
        cnn_section()
    if only == 'cnn':
        return

    for lag, G, regime, tag in ((20, 1, 0, 'C5 sparse lag20 G1'), (20, 1, 2, 'C5 sparse lag20 G1 SORTED rows'),
                                (13, 8, 0, 'C3 sparse lag13 G8'), (10, 2, 1, 'C2 dense lag10 G2')):
        n = K // G if G > 1 else K
        stride = (n + 3) // 4 * 4
        kmers = torch.empty(stride, dtype=torch.int64, device=dev)
        counts = torch.empty((G, 5, stride), dtype=torch.int32, device=dev)
        check(lib.bear_synth_table(ptr(kmers), ptr(counts), stride, 0, n, lag, G, 20, regime, 10, _lib.stream()))
        ws = torch.empty(lib.bear_workspace_doubles(n, lag, 0), dtype=torch.float64, device=dev)
        alpha = torch.tensor([0.1, 1.0, 10.0], dtype=torch.float64, device=dev)
        acc = torch.zeros(G * 3, dtype=torch.float64, device=dev)
        t = timed(lambda: check(lib.bear_bmm_likelihood(ptr(counts), stride, 0, n, G, 5, ptr(alpha), 3, ptr(acc), ptr(ws), _lib.stream())))
        report('bmm_kernel<5> V=3 [%s]' % tag, t, n, 20 * G)
        mat = (torch.randn(lag, 5, 5, dtype=torch.float64, device=dev) * 0.05).contiguous()
        hs = torch.zeros(1, dtype=torch.float64, device=dev)
        flat = torch.zeros(2 + lag * 25, dtype=torch.float64, device=dev)
        for ar in (0, 1):
            t = timed(lambda: check(lib.bear_linear_train_step(ptr(kmers), ptr(counts), stride, 0, n, lag, ptr(mat), ptr(hs), 1.0, ar,
                                                               ptr(flat), None, ptr(ws), _lib.stream())))
            report('linear_train_kernel<%s> [%s]' % ('AR' if ar else 'BEAR', tag), t, n, 28)
        h = torch.ones(1, dtype=torch.float64, device=dev)
        eacc = torch.zeros(11, dtype=torch.float64, device=dev)
        t = timed(lambda: check(lib.bear_eval_step(ptr(kmers), ptr(counts), None, stride, 0, n, lag, _lib.HEAD_LINEAR, ptr(mat), ptr(h), 1,
                                                   ptr(alpha), 3, 7, 0, ptr(eacc), ptr(ws), _lib.stream())))
        report('eval_kernel<LINEAR> H=1 V=3 no-train [%s]' % tag, t, n, 28)
        if G > 1:
            tr = ctypes_off(counts, 5 * stride * 4)
            t = timed(lambda: check(lib.bear_eval_step(ptr(kmers), tr, ptr(counts), stride, 0, n, lag, _lib.HEAD_LINEAR, ptr(mat), ptr(h), 1,
                                                       ptr(alpha), 3, 7, 0, ptr(eacc), ptr(ws), _lib.stream())))
            report('eval_kernel<LINEAR> H=1 V=3 heldout [%s]' % tag, t, n, 48)
        m = min(n, 1 << 24)
        f = torch.full((m, 5), 0.2, dtype=torch.float64, device=dev)
        gf = torch.empty_like(f)
        t = timed(lambda: check(lib.bear_dm_train_step_explicit(ptr(counts), stride, 0, m, ptr(f), ptr(hs), 1.0, 0, ptr(flat), ptr(gf), None,
                                                                ptr(ws), _lib.stream())))
        report('explicit_train_kernel<BEAR> [%s]' % tag, t, m, 100, 'reads 20 B counts + 40 B f, writes 40 B df')
        oh = torch.empty((m, lag, 5), dtype=torch.float64, device=dev)
        t = timed(lambda: check(lib.bear_decode_onehot(ptr(kmers), m, lag, 0, ptr(oh), _lib.stream())))
        report('decode_onehot_kernel [%s]' % tag, t, m, 8 + 40 * lag, 'writes the reference-shaped float64 one-hot')
        dense = torch.empty((m, G, 5), dtype=torch.float64, device=dev)
        t = timed(lambda: check(lib.bear_unpack_counts(ptr(counts), stride, 0, m, G, 5, ptr(dense), _lib.stream())))
        report('unpack_counts_kernel [%s]' % tag, t, m, 60 * G)
        del kmers, counts, oh, dense, f, gf
        torch.cuda.empty_cache()
    with open(os.path.join(ROOT, 'gpurun_out', 'kernel_bench.json'), 'w') as fh:
        json.dump(out, fh, indent=1)


def ctypes_off(t, nbytes):
    import ctypes
    return ctypes.c_void_p(t.data_ptr() + nbytes)


if __name__ == '__main__':
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    main()
