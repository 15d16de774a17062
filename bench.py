#!/usr/bin/env python
"""bench.py -- k-mer transitions/s of the BEAR hot path (train + eval) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--rows R]
    (N > 1: python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...)

Workload (BASELINE.json configs[4], the config the metric is quoted on): linear-AR BEAR, lag 20,
1 group, synthetic table of ~2B distinct k-mers (2^31 rows, sparse counts), sharded by rows over the
N GPUs (strong scaling: the table is fixed, each rank keeps rows/N resident in HBM).  One step = one
training pass (fused fwd+bwd kernel over the shard -> one allreduce of the 502-double flat buffer ->
Adam) plus one evaluation pass (BEAR/AR/BMM likelihoods + accuracies) over the same shard.
value = 2 * rows / step time  (train rows + eval rows, whole job).

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LAG = 20
TRAIN_BYTES_PER_ROW = 28          # 8 B packed k-mer + 5 x 4 B counts (one column)  SURVEY.md 8(d)
EVAL_BYTES_PER_ROW = 28           # ds_loc_train = -1: k-mer + the test column
# dram__bytes_read.sum + dram__bytes_write.sum of linear_train2_kernel per row, from the ncu --set full capture
# profiles/r1_fused_raw.csv (1.880 GB + 0.006 GB over 67 108 864 rows)
TRAIN_DRAM_BYTES_PER_ROW_NCU = 28.10
DEFAULT_ROWS = 1 << 31


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--rows', type=int, default=DEFAULT_ROWS, help='total table rows (all GPUs)')
    ap.add_argument('--cpu-rows', type=int, default=1 << 18, help='rows of the CPU-baseline sample')
    ap.add_argument('--e2e-rows', type=int, default=1 << 26, help='rows per GPU of the host-resident e2e sample')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = 'clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index):
        self.index, self.rows, self.stop = index, [], threading.Event()
        self.t = threading.Thread(target=self.run, daemon=True)

    def run(self):
        while not self.stop.is_set():
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                      '--format=csv,noheader,nounits'], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(',')]
                if len(parts) >= 6:
                    self.rows.append(parts)
            except Exception:
                pass
            self.stop.wait(0.2)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        sm = [float(r[0]) for r in self.rows if r[0].replace('.', '').isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith('active') for r in self.rows)]
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': reasons, 'samples': len(self.rows)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle (op-for-op restatement of the reference's TF graph) on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_step_fn(rows, seed=0):
    """Returns (step callable, rows) running one train pass + one eval pass of the oracle on a
    synthetic sample with the workload's statistics (lag 20, sparse counts, linear head)."""
    import numpy as np
    import torch
    from oracle import bear_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    rng = np.random.default_rng(seed)
    idx = rng.integers(0, 4, size=(rows, LAG))
    onehot = torch.zeros(rows, LAG, 5, dtype=torch.float64)
    onehot.scatter_(2, torch.from_numpy(idx)[..., None], 1.0)
    N = 1 + rng.poisson(2.0, size=rows)
    dom = rng.integers(0, 4, size=rows)
    counts = np.zeros((rows, 5))
    for t in range(int(N.max())):
        live = N > t
        letter = np.where(rng.random(rows) < 0.7, dom, rng.integers(0, 4, size=rows))
        np.add.at(counts, (np.flatnonzero(live), letter[live]), 1.0)
    counts = torch.from_numpy(counts)
    gen = torch.Generator().manual_seed(seed)
    params = O.init_linear(LAG, 4, gen)
    h_signed = torch.zeros((), dtype=torch.float64)
    opt = O.KerasAdam([h_signed] + params, 0.01)
    van = np.array([0.1, 1.0, 10.0])

    def step():
        loss, _, grads = O.train_step_grads(onehot, counts, h_signed, params, 'linear', rows, False)
        opt.apply([h_signed] + params, grads)
        f = O.ar_linear(onehot, params)
        out = O.evaluation([(onehot, f, counts, None)], torch.exp(h_signed), van)
        return float(loss), float(out[0])
    return step


def run_reference(args, rank):
    """--impl reference: the reference's CPU path.  TensorFlow is not installable in this image (no
    wheel in /opt/wheelhouse, no network), so the timed implementation is the oracle port of the
    reference graph (kind = "port"), on all host cores, on a bounded sample of the same workload."""
    if rank != 0:
        return
    rows = args.cpu_rows
    step = cpu_step_fn(rows)
    for _ in range(max(args.warmup, 1)):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    value = 2 * rows / dt
    cores = os.cpu_count() or 1
    sample = '%d synthetic lag-20 rows per step (train pass + eval pass), torch-CPU float64 oracle port' % rows
    print_result(json.dumps({
        'impl': 'reference', 'metric': 'kmer_transitions_per_s_train_plus_eval', 'value': value,
        'unit': 'k-mer transition rows/s', 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': dt * 1e3, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
        'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': 'C5 linear-AR BEAR lag 20, 1 group, sparse synthetic counts (bounded CPU sample)',
                   'rows_per_step': rows, 'lag': LAG},
        'cpu_baseline': {'value': value, 'unit': 'k-mer transition rows/s', 'cores': cores, 'kind': 'port',
                         'sample': sample},
        'e2e': {'value': value, 'unit': 'k-mer transition rows/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }))


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_ours(args, rank, world_size, local_rank):
    import numpy as np
    import torch
    import torch.distributed as dist
    from bear_b200 import _lib
    from bear_b200._lib import lib, check, ptr

    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world_size > 1:
        dist.init_process_group('nccl', device_id=dev)

    rows_total = args.rows
    per = -(-rows_total // world_size)
    row_begin = min(rank * per, rows_total)
    n = min(per, rows_total - row_begin)
    stride = (n + 3) // 4 * 4
    kmers = torch.empty(stride, dtype=torch.int64, device=dev)
    counts = torch.empty((1, 5, stride), dtype=torch.int32, device=dev)
    check(lib.bear_synth_table(ptr(kmers), ptr(counts), stride, row_begin, n, LAG, 1, 20, 0, 10, _lib.stream()))

    P = LAG * 25
    gen = torch.Generator().manual_seed(0)
    mat = torch.randn(LAG, 5, 5, dtype=torch.float64, generator=gen)
    mat = (0.05 * mat / mat.pow(2).sum(1, keepdim=True).sqrt())
    flat_params = torch.zeros(1 + P, dtype=torch.float64, device=dev)
    flat_params[1:] = mat.reshape(-1).to(dev)
    grad = torch.zeros(2 + P, dtype=torch.float64, device=dev)
    m_adam, v = torch.zeros_like(flat_params), torch.zeros_like(flat_params)
    step_ctr = torch.zeros(1, dtype=torch.int64, device=dev)
    ws = torch.empty(lib.bear_workspace_doubles(n, LAG, P), dtype=torch.float64, device=dev)
    hvals = torch.ones(1, dtype=torch.float64, device=dev)
    van = torch.tensor([0.1, 1.0, 10.0], dtype=torch.float64, device=dev)
    acc = torch.zeros(2 + 6 + 3, dtype=torch.float64, device=dev)
    scale = 1.0                        # full-batch: num_kmers / batch rows = 1
    col = ctypes_ptr(counts)
    train_events = []

    def train_pass(k_t, c_ptr, n_rows, pitch, record=False):
        grad.zero_()
        if record:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        check(lib.bear_linear_train_step(ptr(k_t), c_ptr, pitch, 0, n_rows, LAG, ptr(flat_params[1:]),
                                         ptr(flat_params[:1]), scale, 0, ptr(grad), None, ptr(ws), _lib.stream()))
        if record:
            e1.record()
            train_events.append((e0, e1))
        if world_size > 1:
            dist.all_reduce(grad)
        check(lib.bear_adam_update(ptr(flat_params), ptr(grad[1:]), ptr(m_adam), ptr(v), 1 + P, 0.01, 0.9, 0.999, 1e-7,
                                   ptr(step_ctr), _lib.stream()))

    def eval_pass(k_t, c_ptr, n_rows, pitch):
        acc.zero_()
        hvals.copy_(torch.exp(flat_params[:1]))
        check(lib.bear_eval_step(ptr(k_t), c_ptr, None, pitch, 0, n_rows, LAG, _lib.HEAD_LINEAR, ptr(flat_params[1:]),
                                 ptr(hvals), 1, ptr(van), 3, 12345, ptr(acc), ptr(ws), _lib.stream()))
        if world_size > 1:
            dist.all_reduce(acc)

    def step(record=False):
        train_pass(kmers, col, n, stride, record)
        eval_pass(kmers, col, n, stride)

    def barrier():
        if world_size > 1:
            dist.barrier()
        torch.cuda.synchronize()

    w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(args.warmup):
        if i == args.warmup - 1:
            w0.record()                   # the last warm-up step alone: the first ones carry one-time costs (NCCL init)
        step()
    w1.record()
    barrier()
    # nvidia-smi answers in 0.1-0.3 s: when the K timed steps are shorter than ~1 s (small shards at N = 8) the same
    # step keeps running AFTER the timed region has been closed (s1 recorded) so that the sampler sees the clocks
    # under this load; the number of extra steps is the same on every rank (they contain the allreduce)
    est = torch.tensor([w0.elapsed_time(w1) if args.warmup > 0 else 1e9], dtype=torch.float64, device=dev)
    if world_size > 1:
        dist.all_reduce(est, op=dist.ReduceOp.MAX)
    est_ms = max(float(est), 1e-3)
    extra_steps = 0 if args.steps * est_ms >= 1000.0 else min(int((1000.0 - args.steps * est_ms) / est_ms) + 1, 400)
    with ClockSampler(local_rank) as clocks:
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(args.steps):
            step(record=True)
        s1.record()
        for _ in range(extra_steps):
            step()
        barrier()
    ms = torch.tensor([s0.elapsed_time(s1)], dtype=torch.float64, device=dev)
    if world_size > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(ms) / args.steps
    value = 2.0 * rows_total / (ms_per_step * 1e-3)
    train_ms = statistics.mean(a.elapsed_time(b) for a, b in train_events)
    loss_now = float(grad[0])
    results = acc.cpu().tolist()

    # ---- e2e: the same step from HOST-resident (pinned) buffers, copies inside the timed region ----
    # The host side holds the table in the library's compact transfer format (k-mer byte planes, 4-bit count planes
    # + escapes, include/bear_b200.h: 7.6 B per row here instead of 28 B); every step copies it H2D chunk by chunk on a copy
    # stream, expands each chunk on the device (bear_expand_table) and trains on it while the next chunk is in flight.
    import ctypes
    e_rows = min(args.e2e_rows, n)
    e_stride = (e_rows + 3) // 4 * 4
    hk = np.ascontiguousarray(kmers[:e_stride].cpu().numpy())
    hc = np.ascontiguousarray(counts[:, :, :e_stride].cpu().numpy())
    dk = torch.zeros(e_stride, dtype=torch.int64, device=dev)
    dc = torch.zeros((1, 5, e_stride), dtype=torch.int32, device=dev)
    out_host = torch.empty(2 + P + acc.numel(), dtype=torch.float64).pin_memory()

    copy_stream = torch.cuda.Stream(device=dev)
    # the copy of step s+1 runs behind the kernels of step s, so chunks only shorten the un-prefetched first step:
    # two of them cost fewer launches than eight (measured +6 % on the leg)
    n_chunks = int(os.environ.get('BEAR_E2E_CHUNKS', 2))
    bounds = [(e_rows * i // n_chunks) // 4 * 4 for i in range(n_chunks)] + [e_rows]
    chunks = []                       # (lo, rows, pinned compact bytes, device buffers, pinned escapes, device escapes, n_esc)
    for lo, hi in zip(bounds[:-1], bounds[1:]):
        m = hi - lo
        bits = lib.bear_compact_choose_wire(ptr(hk), ptr(hc), e_stride, lo, m, LAG, 0, 1)   # 4-bit counts, start runs as escapes
        check(bits)
        nb = lib.bear_compact_bytes(m, LAG, 0, 1, bits)
        hb = torch.empty(nb, dtype=torch.uint8).pin_memory()
        cap = 1 << 16
        while True:
            esc = np.empty((cap, 3), dtype=np.uint32)
            need = ctypes.c_int64(0)
            check(lib.bear_compact_table(ptr(hk), ptr(hc), e_stride, lo, m, LAG, 0, 1, bits, ctypes.c_void_p(hb.data_ptr()),
                                         ptr(esc), cap, ctypes.byref(need)))
            if need.value <= cap:
                break
            cap = int(need.value)
        he = torch.from_numpy(esc[:need.value].view(np.int32).copy()).pin_memory() if need.value else None
        dbs = [torch.empty_like(hb, device=dev) for _ in range(2)]            # double-buffered on the device
        des = [torch.empty_like(he, device=dev) for _ in range(2)] if he is not None else [None, None]
        chunks.append((lo, m, hb, dbs, he, des, int(need.value), bits))
    del hk, hc
    col_d = ctypes_ptr(dc)
    pending, state = {}, {'k': 0}

    def issue_copies(slot):
        """H2D of one step's input (all chunks) into device buffer set `slot`, on the copy stream."""
        evs = []
        with torch.cuda.stream(copy_stream):
            for lo, m, hb, dbs, he, des, ne, bits in chunks:
                dbs[slot].copy_(hb, non_blocking=True)
                if he is not None:
                    des[slot].copy_(he, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
                evs.append(ev)
        pending[slot] = evs

    def e2e_step():
        # every step copies its whole input H2D and reads its results back D2H; the copy of step s+1 is issued at the
        # start of step s (the input does not depend on the results), so it overlaps the kernels of step s
        main = torch.cuda.current_stream()
        slot = state['k'] & 1
        if slot not in pending:
            issue_copies(slot)
        events = pending.pop(slot)
        issue_copies(slot ^ 1)
        grad.zero_()
        for (lo, m, hb, dbs, he, des, ne, bits), ev in zip(chunks, events):
            main.wait_event(ev)
            check(lib.bear_expand_table(ptr(dbs[slot]), ptr(des[slot]), ne, m, LAG, 0, 1, bits, ptr(dk), ptr(dc), e_stride,
                                        lo, _lib.stream()))
            check(lib.bear_linear_train_step(ptr(dk), col_d, e_stride, lo, m, LAG, ptr(flat_params[1:]),
                                             ptr(flat_params[:1]), scale, 0, ptr(grad), None, ptr(ws), _lib.stream()))
        if world_size > 1:
            dist.all_reduce(grad)
        check(lib.bear_adam_update(ptr(flat_params), ptr(grad[1:]), ptr(m_adam), ptr(v), 1 + P, 0.01, 0.9, 0.999, 1e-7,
                                   ptr(step_ctr), _lib.stream()))
        eval_pass(dk, col_d, e_rows, e_stride)
        out_host[:2 + P].copy_(grad, non_blocking=True)
        out_host[2 + P:].copy_(acc, non_blocking=True)
        main.synchronize()
        state['k'] += 1

    for _ in range(2):
        e2e_step()
    barrier()
    e_steps = max(3, min(args.steps, 10))
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()                       # device clock; every step ends with a stream synchronize (the D2H read)
    for _ in range(e_steps):
        e2e_step()
    g1.record()
    barrier()
    e_dt = torch.tensor([g0.elapsed_time(g1) * 1e-3 / e_steps], dtype=torch.float64, device=dev)
    if world_size > 1:
        dist.all_reduce(e_dt, op=dist.ReduceOp.MAX)
    e2e_value = 2.0 * e_rows * world_size / float(e_dt)
    h2d = sum(c[2].numel() + (c[4].numel() * 4 if c[4] is not None else 0) for c in chunks)
    d2h = out_host.numel() * 8

    if rank != 0:
        if world_size > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as fh:
            peaks = json.load(fh)
    except Exception:
        pass
    peak = float(peaks.get('hbm_gbs', 6650.0))
    achieved = TRAIN_BYTES_PER_ROW * n / (train_ms * 1e-3) / 1e9
    line = {
        'metric': 'kmer_transitions_per_s_train_plus_eval', 'value': value, 'unit': 'k-mer transition rows/s',
        'n_gpus': world_size, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_per_step,
        'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': 'C5 linear-AR BEAR lag 20, 1 group, %d distinct synthetic k-mers (sparse counts), '
                               'row-sharded over %d GPU(s); step = full-shard train pass + eval pass' % (rows_total, world_size),
                   'rows_total': rows_total, 'rows_per_gpu': n, 'lag': LAG, 'groups': 1,
                   'l2': 'inputs larger than L2 (%.1f GB per GPU per pass)' % (TRAIN_BYTES_PER_ROW * n / 1e9)},
        'clocks': dict(clocks.summary(), sampled_over='the %d timed steps + %d identical untimed steps after them'
                       % (args.steps, extra_steps)),
        'e2e': {'value': e2e_value, 'unit': 'k-mer transition rows/s', 'h2d_bytes_per_step': h2d,
                'd2h_bytes_per_step': d2h, 'rows_per_gpu_per_step': e_rows,
                'host_format': 'compact transfer format (k-mer byte planes, %d-bit count planes, escapes%s), %.1f B/row'
                               % (chunks[0][7] & 15, ' incl. start-run lengths' if chunks[0][7] & 16 else '', h2d / e_rows)},
        'gpu_launches': args.steps * 6,       # timed region: train + reduce, adam + bump, eval + reduce per step
        'roofline': {'bound': 'hbm', 'kernel': 'linear_train2_kernel<false>', 'achieved': achieved, 'peak': peak,
                     'unit': 'GB/s', 'frac': achieved / peak, 'traffic': TRAIN_DRAM_BYTES_PER_ROW_NCU * n,
                     'traffic_note': 'bytes per launch = ncu dram read+write per row (profiles/r1_fused_raw.csv) x rows',
                     'algorithmic_bytes': TRAIN_BYTES_PER_ROW * n,
                     'peak_source': 'MEASURED_PEAKS.json hbm_gbs' if 'hbm_gbs' in peaks else 'fallback 6650 GB/s',
                     'kernel_ms': train_ms, 'bytes_per_row': TRAIN_BYTES_PER_ROW},
        'train_rows_per_s': n * world_size / (train_ms * 1e-3),
        'check': {'loss': loss_now, 'eval_acc': results},
    }
    if world_size == 1 and not args.no_cpu_baseline:
        stepf = cpu_step_fn(args.cpu_rows)
        stepf()
        t0 = time.perf_counter()
        reps = 0
        while reps < 3 or (time.perf_counter() - t0 < 10 and reps < 50):
            stepf()
            reps += 1
        dt = (time.perf_counter() - t0) / reps
        line['cpu_baseline'] = {'value': 2 * args.cpu_rows / dt, 'unit': 'k-mer transition rows/s',
                                'cores': os.cpu_count() or 1, 'kind': 'port',
                                'sample': '%d synthetic lag-20 rows x %d steps (train pass + eval pass), torch-CPU float64 '
                                          'oracle port of the reference TF graph' % (args.cpu_rows, reps)}
    print_result(json.dumps(line))
    if world_size > 1:
        dist.destroy_process_group()


_RESULT_OUT = None               # the process's original stdout, saved by main() before fd 1 is pointed at stderr


def print_result(line):
    out = _RESULT_OUT or sys.stdout
    out.write(line + '\n')
    out.flush()


def ctypes_ptr(t):
    import ctypes
    return ctypes.c_void_p(t.data_ptr())


def main():
    args = parse()
    # stdout carries exactly ONE line, the JSON result: libraries that print to file descriptor 1 (NCCL writes
    # "NCCL version ..." there on communicator creation) are sent to stderr for the rest of the run
    global _RESULT_OUT
    sys.stdout.flush()
    _RESULT_OUT = os.fdopen(os.dup(1), 'w')
    os.dup2(2, 1)
    rank = int(os.environ.get('RANK', 0))
    world_size = int(os.environ.get('WORLD_SIZE', 1))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    if args.impl == 'reference':
        return run_reference(args, rank)
    from bear_b200 import build
    if rank == 0:
        build.build()
    run_ours(args, rank, world_size, local_rank)


if __name__ == '__main__':
    main()
