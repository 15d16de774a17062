#!/usr/bin/env python
"""bench.py -- k-mer transitions/s of the BEAR hot path (train + eval) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--rows R] [--batch-rows B]
    (N > 1: python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...)

Headline workload (BASELINE.json configs[4], the config the metric is quoted on): linear-AR BEAR, lag 20,
1 group, synthetic table of ~2B distinct k-mers (2^31 rows, sparse counts), sharded by rows over the N GPUs
(strong scaling: the table is fixed, each rank keeps rows/N resident in HBM).  One step = one training pass
over the table (per optimizer step: fused fwd+bwd kernel over this rank's part of the batch -> ONE allreduce of
the 502-double flat buffer -> one optimizer launch) plus one evaluation pass (BEAR/AR/BMM likelihoods +
accuracies).  --batch-rows B sets the GLOBAL batch of an optimizer step (default: the whole table, one optimizer
step per pass); with more than one batch per pass the pass is replayed from a CUDA graph (kernels + NCCL).
value = 2 * rows / step time  (train rows + eval rows, whole job).

Extra legs in the same JSON line: `e2e` (host-resident input, copies inside the timed region), `batch_scaling`
(the same step at global batches of 2^22 and 2^26 rows), `probe` (N-rank result == 1-rank result on a fixed
2^20-row probe table), `configs` (BASELINE.json configs[0..3] through the public API, each with its roofline
fraction and CPU baseline), `cpu_baseline`.  Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement".
"""
import argparse
import ctypes
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
# dram__bytes_read.sum + dram__bytes_write.sum of linear_train_tc_kernel per row, from the ncu --set full capture
# profiles/r2_train_tc_raw.csv (1.8796 GB + 0.0093 GB over 67 108 864 rows)
TRAIN_DRAM_BYTES_PER_ROW_NCU = 28.15
DEFAULT_ROWS = 1 << 31
TRAIN_KERNEL = 'linear_train_tc_kernel<false, 5>'
METRIC = 'kmer_transitions_per_s_train_plus_eval'
UNIT = 'k-mer transition rows/s'


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--rows', type=int, default=DEFAULT_ROWS, help='total table rows (all GPUs)')
    ap.add_argument('--batch-rows', type=int, default=0, help='global rows per optimizer step (0: the whole table)')
    ap.add_argument('--cpu-rows', type=int, default=1 << 22, help='rows of the CPU-baseline sample')
    ap.add_argument('--e2e-rows', type=int, default=1 << 26, help='rows per GPU of the host-resident e2e sample')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-extra-legs', action='store_true', help='headline + e2e only (no batch sweep, configs, probe)')
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


def bind_to_gpu_numa(index):
    """Pin this process (and with it the pages of the pinned staging buffers it allocates afterwards) to the NUMA node
    the GPU hangs off: eight ranks copying H2D at once otherwise share one socket's memory controllers and PCIe root."""
    try:
        bus = subprocess.run(['nvidia-smi', '-i', str(index), '--query-gpu=pci.bus_id', '--format=csv,noheader'],
                             capture_output=True, text=True, timeout=10).stdout.strip().lower()
        bus = bus[-12:] if len(bus) >= 12 else bus                    # 00000000:1B:00.0 -> 0000:1b:00.0
        node = int(open('/sys/bus/pci/devices/%s/numa_node' % bus).read())
        if node < 0:
            return {'numa_node': None}
        cpus = set()
        for part in open('/sys/devices/system/node/node%d/cpulist' % node).read().strip().split(','):
            lo, _, hi = part.partition('-')
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return {'numa_node': node, 'cpus': len(cpus)}
    except Exception as e:                                            # containers without sysfs / nvidia-smi
        return {'numa_node': None, 'note': type(e).__name__}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle (op-for-op restatement of the reference's TF graph) on the host cores
# ------------------------------------------------------------------------------------------------
def synth_cpu_rows(rows, lag, seed=0, dense=False):
    """k-mer byte strings ('S<lag>') and counts [rows, 5] with the workload's statistics."""
    import numpy as np
    rng = np.random.default_rng(seed)
    idx = rng.integers(0, 4, size=(rows, lag), dtype=np.uint8)
    kmers = np.frombuffer(b'ACGT', dtype=np.uint8)[idx].copy().view('S%d' % lag).reshape(rows)
    if dense:
        N = np.maximum(1, np.round(rng.lognormal(np.log(300.0), 1.5, size=rows))).astype(np.int64)
        p = rng.dirichlet(0.3 * np.ones(4), size=rows)
        counts = np.zeros((rows, 5))
        counts[:, :4] = np.floor(N[:, None] * p)
    else:
        N = 1 + rng.poisson(2.0, size=rows)
        dom = rng.integers(0, 4, size=rows)
        counts = np.zeros((rows, 5))
        for t in range(int(N.max())):
            live = N > t
            letter = np.where(rng.random(rows) < 0.7, dom, rng.integers(0, 4, size=rows))
            np.add.at(counts, (np.flatnonzero(live), letter[live]), 1.0)
    return kmers, counts


def cpu_step_fn(rows, lag=LAG, head='linear', seed=0, dense=False, ref=False):
    """Returns a step callable: one train pass + one eval pass of the oracle on a synthetic sample with the workload's
    statistics, from the k-mer byte strings (the one-hot input is rebuilt every step, as the reference's tf.data
    map does, bear_net.py:268-273)."""
    import numpy as np
    import torch
    from oracle import bear_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    kmers, counts = synth_cpu_rows(rows, lag, seed, dense)
    counts = torch.from_numpy(counts)
    gen = torch.Generator().manual_seed(seed)
    params = O.init_cnn(lag, 4, gen, filter_width=3) if head == 'cnn' else O.init_linear(lag, 4, gen)
    h_signed = torch.zeros((), dtype=torch.float64)
    opt = O.KerasAdam([h_signed] + params, 0.01)
    van = np.array([0.1, 1.0, 10.0])
    refc = torch.from_numpy(np.random.default_rng(seed + 1).poisson(0.3, size=(rows, 5)).astype(np.float64)) if ref else None

    def step():
        onehot = O.one_hot_bytes(kmers)
        if ref:      # bear_ref evaluation only (C2): stop net, Jukes-Cantor mix of the reference column
            f = O.ar_ref(onehot, O.ref_counts_map(refc, 4), torch.tensor(np.log(1 / 30)), torch.tensor(-np.log(100.0)),
                         lambda x: O.ar_stop(x, 4), 4)
            out = O.evaluation([(onehot, f, counts, None)], torch.exp(h_signed), van)
            return 0.0, float(out[0])
        loss, _, grads = O.train_step_grads(onehot, counts, h_signed, params, head, rows, False)
        opt.apply([h_signed] + params, grads)
        f = O.ar_cnn(onehot, params) if head == 'cnn' else O.ar_linear(onehot, params)
        out = O.evaluation([(onehot, f, counts, None)], torch.exp(h_signed), van)
        return float(loss), float(out[0])
    return step


def time_cpu(step, min_reps=2, budget_s=12.0, max_reps=50):
    step()
    t0 = time.perf_counter()
    reps = 0
    while reps < min_reps or (time.perf_counter() - t0 < budget_s and reps < max_reps):
        step()
        reps += 1
    return (time.perf_counter() - t0) / reps, reps


def run_reference(args, rank):
    """--impl reference: the reference's CPU path.  TensorFlow is not installable in this image (no wheel in
    /opt/wheelhouse, no network), so the timed implementation is the oracle port of the reference graph
    (kind = "port"), on all host cores, on a bounded sample of the same workload."""
    if rank != 0:
        return
    rows = args.cpu_rows
    step = cpu_step_fn(rows)
    for _ in range(max(min(args.warmup, 2), 1)):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    value = 2 * rows / dt
    cores = os.cpu_count() or 1
    sample = ('%d synthetic lag-20 rows per step (k-mer strings -> one-hot -> train pass + eval pass), torch-CPU float64 '
              'oracle port of the reference TF graph' % rows)
    print_result(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': dt * 1e3, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
        'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': 'C5 linear-AR BEAR lag 20, 1 group, sparse synthetic counts (bounded CPU sample)',
                   'rows_per_step': rows, 'lag': LAG},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }))


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
class Ctx:
    pass


def make_ctx(rank, world_size, local_rank):
    import torch
    import torch.distributed as dist
    from bear_b200 import _lib
    c = Ctx()
    c.torch, c.dist, c._lib, c.lib, c.check, c.ptr = torch, dist, _lib, _lib.lib, _lib.check, _lib.ptr
    c.rank, c.world, c.local_rank = rank, world_size, local_rank
    torch.cuda.set_device(local_rank)
    c.dev = torch.device('cuda', local_rank)
    if world_size > 1:
        dist.init_process_group('nccl', device_id=c.dev)
    return c


def allreduce(c, t, op=None):
    if c.world > 1:
        c.dist.all_reduce(t, op=op or c.dist.ReduceOp.SUM)
    return t


def barrier(c):
    if c.world > 1:
        c.dist.barrier()
    c.torch.cuda.synchronize()


def max_over_ranks(c, x):
    t = c.torch.tensor([float(x)], dtype=c.torch.float64, device=c.dev)
    return float(allreduce(c, t, c.dist.ReduceOp.MAX) if c.world > 1 else t)


def synth_table(c, n, lag, G, seed, regime, row_begin=0):
    torch = c.torch
    stride = max((n + 3) // 4 * 4, 4)
    kmers = torch.empty(stride, dtype=torch.int64, device=c.dev)
    counts = torch.empty((G, 5, stride), dtype=torch.int32, device=c.dev)
    c.check(c.lib.bear_synth_table(c.ptr(kmers), c.ptr(counts), stride, row_begin, n, lag, G, seed, regime, 10, c._lib.stream()))
    return kmers, counts, stride


def shard_range(rows_total, rank, world):
    per = -(-rows_total // world)
    lo = min(rank * per, rows_total)
    return lo, min(per, rows_total - lo)


class LinearModel:
    """Flat parameters + optimizer state of the linear-head BEAR model, and its train / eval passes over a resident table
    through the C-ABI (the calls bear_net.train / bear_net.evaluation make)."""

    def __init__(self, c, lag, seed=0):
        torch = c.torch
        self.c, self.lag, self.P = c, lag, lag * 25
        gen = torch.Generator().manual_seed(seed)
        mat = torch.randn(lag, 5, 5, dtype=torch.float64, generator=gen)
        mat = 0.05 * mat / mat.pow(2).sum(1, keepdim=True).sqrt()
        self.flat = torch.zeros(1 + self.P, dtype=torch.float64, device=c.dev)
        self.flat[1:] = mat.reshape(-1).to(c.dev)
        self.init = self.flat.clone()
        self.grad = torch.zeros(2 + self.P, dtype=torch.float64, device=c.dev)
        self.m, self.v = torch.zeros_like(self.flat), torch.zeros_like(self.flat)
        self.step_ctr = torch.zeros(1, dtype=torch.int64, device=c.dev)
        self.loss = torch.zeros(1, dtype=torch.float64, device=c.dev)
        self.hvals = torch.ones(1, dtype=torch.float64, device=c.dev)
        self.van = torch.tensor([0.1, 1.0, 10.0], dtype=torch.float64, device=c.dev)
        self.acc = torch.zeros(2 + 6 + 3, dtype=torch.float64, device=c.dev)
        self.ws = torch.empty(c.lib.bear_workspace_doubles(0, lag, self.P), dtype=torch.float64, device=c.dev)
        self.launches = 0

    def reset(self):
        self.flat.copy_(self.init)
        self.m.zero_()
        self.v.zero_()
        self.step_ctr.zero_()
        self.grad.zero_()

    def train_kernel(self, kmers, col, stride, r0, n, scale):
        c = self.c
        c.check(c.lib.bear_linear_train_step(c.ptr(kmers), col, stride, r0, n, self.lag, c.ptr(self.flat[1:]), c.ptr(self.flat[:1]),
                                             scale, 0, c.ptr(self.grad), None, c.ptr(self.ws), c._lib.stream()))
        self.launches += 2

    def optimizer(self, lr=0.01):
        c = self.c
        allreduce(c, self.grad)
        c.check(c.lib.bear_adam_step(c.ptr(self.flat), c.ptr(self.grad), c.ptr(self.m), c.ptr(self.v), 1 + self.P, lr, 0.9, 0.999,
                                     1e-7, c.ptr(self.step_ctr), c.ptr(self.loss), 1.0, 1, c._lib.stream()))
        self.launches += 1

    def train_pass(self, kmers, col, stride, ranges, scale):
        for r0, n in ranges:
            self.train_kernel(kmers, col, stride, r0, n, scale)
            self.optimizer()

    def eval_pass(self, kmers, col, stride, n, row_id0, seed=12345):
        c = self.c
        self.acc.zero_()
        c.torch.exp(self.flat[:1], out=self.hvals)
        c.check(c.lib.bear_eval_step(c.ptr(kmers), col, None, stride, 0, n, self.lag, c._lib.HEAD_LINEAR, c.ptr(self.flat[1:]),
                                     c.ptr(self.hvals), 1, c.ptr(self.van), 3, seed, row_id0, c.ptr(self.acc), c.ptr(self.ws),
                                     c._lib.stream()))
        allreduce(c, self.acc)
        self.launches += 2


def batch_ranges(n_local, rows_total, batch_rows, world):
    """Local (row0, n) of every optimizer step and the loss scale num_kmers / global batch (bear_net.py:190)."""
    if batch_rows <= 0 or batch_rows >= rows_total:
        return [(0, n_local)], 1.0
    per = max(batch_rows // world, 1)
    return [(r0, min(per, n_local - r0)) for r0 in range(0, n_local, per)], float(rows_total) / float(per * world)


def timed_steps(c, step, steps, warmup, graph=False, after=None):
    """`warmup` untimed steps, then `steps` timed ones between events (barrier + synchronize on both sides); returns
    ms per step (max over ranks).  graph: the step is captured into a CUDA graph (kernels + NCCL) and replayed."""
    torch = c.torch
    run = step
    if graph:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            step()                                 # eager once on a side stream (warms up what capture needs)
        torch.cuda.current_stream().wait_stream(side)
        barrier(c)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            step()
        run = g.replay
    for _ in range(warmup):
        run()
    barrier(c)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        run()
    e1.record()
    if after is not None:
        after()
    barrier(c)
    return max_over_ranks(c, e0.elapsed_time(e1) / steps)


def run_ours(args, c):
    torch, lib, check, ptr, _lib = c.torch, c.lib, c.check, c.ptr, c._lib
    import numpy as np
    rank, world = c.rank, c.world
    numa = bind_to_gpu_numa(c.local_rank)
    rows_total = args.rows
    row_begin, n = shard_range(rows_total, rank, world)
    kmers, counts, stride = synth_table(c, n, LAG, 1, 20, 0, row_begin)
    col = ctypes.c_void_p(counts.data_ptr())
    model = LinearModel(c, LAG)
    ranges, scale = batch_ranges(n, rows_total, args.batch_rows, world)
    nb = len(ranges)
    train_events = []

    def step(record=False):
        if record and nb == 1:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            model.train_kernel(kmers, col, stride, 0, n, scale)
            e1.record()
            train_events.append((e0, e1))
            model.optimizer()
        else:
            model.train_pass(kmers, col, stride, ranges, scale)
        model.eval_pass(kmers, col, stride, n, row_begin)

    use_graph = nb > 4
    run = step
    if use_graph:                                  # launch-bound: replay the whole step (kernels + NCCL) from a CUDA graph
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            step()
        torch.cuda.current_stream().wait_stream(side)
        barrier(c)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            step()
        run = graph.replay
    w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(args.warmup):
        if i == args.warmup - 1:
            w0.record()                   # the last warm-up step alone: the first ones carry one-time costs (NCCL init)
        run()
    w1.record()
    barrier(c)
    # nvidia-smi answers in 0.1-0.3 s: when the K timed steps are shorter than ~1 s (small shards at N = 8) the same
    # step keeps running AFTER the timed region has been closed (s1 recorded) so that the sampler sees the clocks
    # under this load; the number of extra steps is the same on every rank (they contain the allreduce)
    est_ms = max(max_over_ranks(c, w0.elapsed_time(w1) if args.warmup > 0 else 1e9), 1e-3)
    extra_steps = 0 if args.steps * est_ms >= 1000.0 else min(int((1000.0 - args.steps * est_ms) / est_ms) + 1, 400)
    chk_loss, chk_acc = torch.zeros(1, dtype=torch.float64, device=c.dev), torch.zeros_like(model.acc)
    model.launches = 0
    with ClockSampler(c.local_rank) as clocks:
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(args.steps):
            if use_graph:
                run()
            else:
                step(record=True)
        s1.record()
        chk_loss.copy_(model.loss)        # the results of the LAST TIMED step, read before the extra steps run
        chk_acc.copy_(model.acc)
        launches = model.launches if not use_graph else args.steps * (nb * 3 + 2)
        for _ in range(extra_steps):
            run()
        barrier(c)
    ms_per_step = max_over_ranks(c, s0.elapsed_time(s1) / args.steps)
    value = 2.0 * rows_total / (ms_per_step * 1e-3)
    if train_events:
        train_ms = statistics.mean(a.elapsed_time(b) for a, b in train_events)
        roof_note = 'CUDA events around the kernel inside the timed steps'
    else:          # batched / graph mode: the kernel over the whole shard, timed alone right after the timed steps
        train_ms = timed_steps(c, lambda: model.train_kernel(kmers, col, stride, 0, n, 1.0), 3, 1)
        model.grad.zero_()
        roof_note = 'CUDA events around one whole-shard launch of the kernel, right after the timed steps'
    loss_now, results = float(chk_loss), chk_acc.cpu().tolist()

    extra = {}
    if not args.no_extra_legs:
        extra['batch_scaling'] = leg_batch_scaling(c, args, model, kmers, col, stride, n, rows_total, row_begin)
        extra['probe'] = leg_probe(c)
    e2e = leg_e2e(c, args, model, kmers, counts, n)
    del kmers, counts
    torch.cuda.empty_cache()
    if not args.no_extra_legs:
        extra['configs'] = leg_configs(c, args)

    if rank != 0:
        if world > 1:
            c.dist.destroy_process_group()
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
        'metric': METRIC, 'value': value, 'unit': UNIT,
        'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_per_step,
        'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': 'C5 linear-AR BEAR lag 20, 1 group, %d distinct synthetic k-mers (sparse counts), '
                               'row-sharded over %d GPU(s); step = train pass (%d optimizer step(s) of %d global rows) + eval pass'
                               % (rows_total, world, nb, rows_total if nb == 1 else args.batch_rows),
                   'rows_total': rows_total, 'rows_per_gpu': n, 'lag': LAG, 'groups': 1,
                   'global_batch_rows': rows_total if nb == 1 else args.batch_rows, 'optimizer_steps_per_pass': nb,
                   'cuda_graph': bool(use_graph),
                   'l2': 'inputs larger than L2 (%.1f GB per GPU per pass)' % (TRAIN_BYTES_PER_ROW * n / 1e9)},
        'clocks': dict(clocks.summary(), sampled_over='the %d timed steps + %d identical untimed steps after them'
                       % (args.steps, extra_steps)),
        'e2e': e2e,
        'gpu_launches': launches,         # timed region, per rank: (train + reduce + optimizer) per batch, eval + reduce per pass
        'roofline': {'bound': 'hbm', 'kernel': TRAIN_KERNEL, 'achieved': achieved, 'peak': peak,
                     'unit': 'GB/s', 'frac': achieved / peak, 'traffic': TRAIN_DRAM_BYTES_PER_ROW_NCU * n,
                     'traffic_note': 'bytes per launch = ncu dram read+write per row (profiles/r2_train_tc_raw.csv) x rows',
                     'algorithmic_bytes': TRAIN_BYTES_PER_ROW * n,
                     'peak_source': 'MEASURED_PEAKS.json hbm_gbs' if 'hbm_gbs' in peaks else 'fallback 6650 GB/s',
                     'kernel_ms': train_ms, 'bytes_per_row': TRAIN_BYTES_PER_ROW, 'timing': roof_note},
        'train_rows_per_s': n * world / (train_ms * 1e-3),
        'check': {'loss': loss_now, 'eval_acc': results, 'read': 'after the last timed step, before any untimed step'},
        'numa': numa,
    }
    line.update(extra)
    if world == 1 and not args.no_cpu_baseline:
        dt, reps = time_cpu(cpu_step_fn(args.cpu_rows))
        line['cpu_baseline'] = {'value': 2 * args.cpu_rows / dt, 'unit': UNIT, 'cores': os.cpu_count() or 1, 'kind': 'port',
                                'sample': '%d synthetic lag-20 rows x %d steps (k-mer strings -> one-hot -> train pass + eval pass), '
                                          'torch-CPU float64 oracle port of the reference TF graph' % (args.cpu_rows, reps)}
    print_result(json.dumps(line))
    if world > 1:
        c.dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
# extra legs
# ------------------------------------------------------------------------------------------------
def leg_batch_scaling(c, args, model, kmers, col, stride, n, rows_total, row_begin):
    """The headline step at smaller GLOBAL batches: many optimizer steps per pass, each with its allreduce; the pass is
    replayed from a CUDA graph (kernels + NCCL + optimizer launch).  The driver's 1 -> 8 efficiency uses `value` of the
    main leg; these rows give the same ratio at the batch sizes where the per-step fixed cost matters."""
    out = []
    for b in (1 << 22, 1 << 26):
        if b >= rows_total:
            continue
        ranges, scale = batch_ranges(n, rows_total, b, c.world)
        if len(ranges) > 2048:                       # bound the graph: a slice of the shard with the same batch size
            ranges = ranges[:2048]
        rows_pass = sum(m for _, m in ranges)
        model.reset()

        def step():
            model.train_pass(kmers, col, stride, ranges, scale)
        try:
            ms = timed_steps(c, step, 3, 1, graph=True)
            mode = 'cuda graph (kernels + NCCL allreduce + optimizer)'
        except Exception as e:                       # capture refused (e.g. an NCCL build that cannot be captured)
            c.torch.cuda.synchronize()
            ms = timed_steps(c, step, 3, 1, graph=False)
            mode = 'eager (%s)' % type(e).__name__
        rows_all = allreduce(c, c.torch.tensor([float(rows_pass)], dtype=c.torch.float64, device=c.dev))
        out.append({'global_batch_rows': b, 'optimizer_steps': len(ranges), 'train_rows_per_s': float(rows_all) / (ms * 1e-3),
                    'us_per_optimizer_step': ms * 1e3 / len(ranges), 'mode': mode})
    model.reset()
    return out


def leg_probe(c):
    """N-rank == 1-rank on a fixed 2^20-row probe table: every rank computes its 1/N slice, the flat buffers are
    allreduced, and rank 0 recomputes the whole probe alone.  Loss / gradient / likelihoods to 1e-10 (different summation
    order), integer accuracy counts exactly (the tie-break noise is keyed on the global row)."""
    torch, lib, check, ptr, _lib = c.torch, c.lib, c.check, c.ptr, c._lib
    K = 1 << 20
    lo, m = shard_range(K, c.rank, c.world)
    model = LinearModel(c, LAG, seed=3)
    alpha = torch.tensor([0.1, 1.0, 10.0], dtype=torch.float64, device=c.dev)

    def compute(row0, rows):
        kmers, counts, stride = synth_table(c, rows, LAG, 1, 77, 0, row0)
        col = ctypes.c_void_p(counts.data_ptr())
        model.grad.zero_()
        model.train_kernel(kmers, col, stride, 0, rows, 1.0)
        model.acc.zero_()
        torch.exp(model.flat[:1], out=model.hvals)
        check(lib.bear_eval_step(ptr(kmers), col, None, stride, 0, rows, LAG, _lib.HEAD_LINEAR, ptr(model.flat[1:]),
                                 ptr(model.hvals), 1, ptr(model.van), 3, 4242, row0, ptr(model.acc), ptr(model.ws), _lib.stream()))
        bmm = torch.zeros(3, dtype=torch.float64, device=c.dev)
        check(lib.bear_bmm_likelihood(ptr(counts), stride, 0, rows, 1, 5, ptr(alpha), 3, ptr(bmm), ptr(model.ws), _lib.stream()))
        return torch.cat([model.grad, model.acc, bmm]).clone()

    dist_res = allreduce(c, compute(lo, m))
    res = {'rows': K, 'ranks': c.world}
    if c.rank == 0:
        full = compute(0, K)
        P = 2 + model.P
        d = (dist_res - full).abs()
        res['loss_rel'] = float(d[0] / full[0].abs())
        res['grad_rel_to_largest'] = float(d[1:P].max() / full[1:P].abs().max())
        res['ll_rel'] = float((d[P:P + 5] / full[P:P + 5].abs()).max())                  # ll_ear, ll_arm, ll_van[3]
        res['bmm_rel'] = float((d[-3:] / full[-3:].abs()).max())
        res['accuracy_counts_equal'] = bool(torch.equal(dist_res[P + 5:P + 11], full[P + 5:P + 11]))
        res['ok'] = bool(res['loss_rel'] <= 1e-10 and res['grad_rel_to_largest'] <= 1e-8 and res['ll_rel'] <= 1e-10
                         and res['bmm_rel'] <= 1e-10 and res['accuracy_counts_equal'])
        assert res['ok'], 'N-rank result differs from the 1-rank result on the probe table: %r' % (res,)
    barrier(c)
    return res


def leg_e2e(c, args, model, kmers, counts, n):
    """The headline step from HOST-resident (pinned) buffers, copies inside the timed region.  The host side holds the
    table in the library's compact transfer format (k-mer byte planes, 4-bit count planes + escapes, include/bear_b200.h:
    7.6 B per row here instead of 28 B); every step copies it H2D chunk by chunk on a copy stream, expands each chunk on
    the device (bear_expand_table) and trains on it while the next chunk is in flight.  Also measured: the box's H2D
    ceiling with every rank copying at once, and the host-side compaction rate."""
    torch, lib, check, ptr, _lib = c.torch, c.lib, c.check, c.ptr, c._lib
    import numpy as np
    dev, world = c.dev, c.world
    P = model.P
    model.reset()
    e_rows = min(args.e2e_rows, n)
    e_stride = (e_rows + 3) // 4 * 4
    hk = np.ascontiguousarray(kmers[:e_stride].cpu().numpy())
    hc = np.ascontiguousarray(counts[:, :, :e_stride].cpu().numpy())
    dk = torch.zeros(e_stride, dtype=torch.int64, device=dev)
    dc = torch.zeros((1, 5, e_stride), dtype=torch.int32, device=dev)
    out_host = torch.empty(1 + model.grad.numel() + model.acc.numel(), dtype=torch.float64).pin_memory()
    copy_stream = torch.cuda.Stream(device=dev)
    n_chunks = int(os.environ.get('BEAR_E2E_CHUNKS', 2))
    bounds = [(e_rows * i // n_chunks) // 4 * 4 for i in range(n_chunks)] + [e_rows]
    chunks = []                       # (lo, rows, pinned compact bytes, device buffers, pinned escapes, device escapes, n_esc, wire)
    t_compact = 0.0
    for lo, hi in zip(bounds[:-1], bounds[1:]):
        m = hi - lo
        t0 = time.perf_counter()
        bits = lib.bear_compact_choose_wire(ptr(hk), ptr(hc), e_stride, lo, m, LAG, 0, 1)   # 12-bit count-vector ranks, start runs as escapes
        check(bits)
        nbytes = lib.bear_compact_bytes(m, LAG, 0, 1, bits)
        hb = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
        cap = 1 << 16
        while True:
            esc = np.empty((cap, 3), dtype=np.uint32)
            need = ctypes.c_int64(0)
            check(lib.bear_compact_table(ptr(hk), ptr(hc), e_stride, lo, m, LAG, 0, 1, bits, ctypes.c_void_p(hb.data_ptr()),
                                         ptr(esc), cap, ctypes.byref(need)))
            if need.value <= cap:
                break
            cap = int(need.value)
        t_compact += time.perf_counter() - t0
        he = torch.from_numpy(esc[:need.value].view(np.int32).copy()).pin_memory() if need.value else None
        dbs = [torch.empty_like(hb, device=dev) for _ in range(2)]            # double-buffered on the device
        des = [torch.empty_like(he, device=dev) for _ in range(2)] if he is not None else [None, None]
        chunks.append((lo, m, hb, dbs, he, des, int(need.value), bits))
    del hk, hc
    col_d = ctypes.c_void_p(dc.data_ptr())
    pending, state = {}, {'k': 0}

    def issue_copies(slot):
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
        for (lo, m, hb, dbs, he, des, ne, bits), ev in zip(chunks, events):
            main.wait_event(ev)
            check(lib.bear_expand_table(ptr(dbs[slot]), ptr(des[slot]), ne, m, LAG, 0, 1, bits, ptr(dk), ptr(dc), e_stride,
                                        lo, _lib.stream()))
            model.train_kernel(dk, col_d, e_stride, lo, m, 1.0)
        model.optimizer()
        model.eval_pass(dk, col_d, e_stride, e_rows, 0)
        out_host[:1].copy_(model.loss, non_blocking=True)
        out_host[1:1 + model.acc.numel()].copy_(model.acc, non_blocking=True)
        out_host[1 + model.acc.numel():].copy_(model.grad, non_blocking=True)
        main.synchronize()
        state['k'] += 1

    for _ in range(2):
        e2e_step()
    barrier(c)
    e_steps = max(3, min(args.steps, 10))
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()                       # device clock; every step ends with a stream synchronize (the D2H read)
    for _ in range(e_steps):
        e2e_step()
    g1.record()
    barrier(c)
    e_dt = max_over_ranks(c, g0.elapsed_time(g1) * 1e-3 / e_steps)
    h2d = sum(ch[2].numel() + (ch[4].numel() * 4 if ch[4] is not None else 0) for ch in chunks)
    # the box's H2D ceiling: every rank copies its pinned compact buffers back to back at the same time, nothing else running
    pending.clear()
    barrier(c)
    reps = 6
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(copy_stream):
        c0.record(copy_stream)
        for _ in range(reps):
            for lo, m, hb, dbs, he, des, ne, bits in chunks:
                dbs[0].copy_(hb, non_blocking=True)
        c1.record(copy_stream)
    barrier(c)
    copy_s = max_over_ranks(c, c0.elapsed_time(c1) * 1e-3 / reps)
    per_rank_gbs = sum(ch[2].numel() for ch in chunks) / copy_s / 1e9
    ceiling_rows = e_rows * world / copy_s           # rows/s if a step were nothing but its H2D copy
    return {'value': 2.0 * e_rows * world / e_dt, 'unit': UNIT, 'h2d_bytes_per_step': h2d,
            'd2h_bytes_per_step': out_host.numel() * 8, 'rows_per_gpu_per_step': e_rows,
            'host_format': 'compact transfer format (k-mer byte planes, %s, escapes%s), %.1f B/row'
                           % ('12-bit rank of each count vector' if chunks[0][7] & 15 == 12 else '%d-bit count planes' % (chunks[0][7] & 15),
                              ' incl. start-run lengths' if chunks[0][7] & 16 else '', h2d / e_rows),
            'h2d_ceiling': {'all_ranks_concurrent_gbs_per_rank': per_rank_gbs, 'aggregate_gbs': per_rank_gbs * world,
                            'step_if_copy_only_rows_per_s': 2.0 * ceiling_rows,
                            'e2e_frac_of_copy_only': (2.0 * e_rows * world / e_dt) / (2.0 * ceiling_rows)},
            'host_compaction_rows_per_s': e_rows / t_compact, 'host_compaction_threads': os.cpu_count() or 1}


def leg_configs(c, args):
    """BASELINE.json configs[0..3] through the public API (bear_net / bear_ref / get_var_probs), sharded by rows over the
    ranks like the headline config; every entry carries rows/s, the roofline that bounds it and (rank 0, N = 1) the CPU
    baseline of the oracle port on a bounded sample."""
    torch, lib, check, ptr, _lib = c.torch, c.lib, c.check, c.ptr, c._lib
    import numpy as np
    from bear_b200 import ar_funcs, bear_net, bear_ref, dataloader as dl, get_var_probs as gvp
    peak = 6650.0
    try:
        peak = float(json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'])
    except Exception:
        pass
    world, rank = c.world, c.rank
    out = {}
    cpu = world == 1 and not args.no_cpu_baseline

    def dataset(table, n_local, batch_global, rows_total, row_begin):
        per = max(batch_global // world, 1)
        ranges = [(r0, min(per, n_local - r0)) for r0 in range(0, n_local, per)]
        return dl.KmerDataset(table, per, 1, ranges, [min(batch_global, rows_total)] * len(ranges), None,
                              [row_begin + r0 for r0, _ in ranges])

    def timed(fn, reps=1, warm=1):
        for _ in range(warm):
            fn()
        barrier(c)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        barrier(c)
        return max_over_ranks(c, e0.elapsed_time(e1) * 1e-3 / reps)

    # ---- C1: bundled ysd1 lag-5 table, end to end from the TSV (parse -> upload -> 10 000 Adam steps -> evaluation);
    #      under torchrun every rank parses the file and keeps its slice of the one 1365-row batch (dataloader shards)
    path = os.path.join(ROOT, 'bear_b200', 'data', 'ysd1_lag_5_file_0_preshuf.tsv')
    epochs = 10000

    def c1():
        data = dl.dataloader(path, 'dna', 1500, 3)
        K = 1365
        torch.manual_seed(10)
        params, h_signed, ar_func = bear_net.train(data.repeat(epochs), K, epochs, 0, 'dna', 5, ar_funcs.make_ar_func_linear, {},
                                                   0.01, 'Adam', False)
        ev = bear_net.evaluation(data, 0, 1, 'dna', torch.exp(h_signed), ar_func, np.array([0.1, 1.0, 10.0]), seed=1)
        torch.cuda.synchronize()
        return K, float(ev[3])
    c1()
    barrier(c)
    t0 = time.perf_counter()
    K, perp = c1()
    dt = max_over_ranks(c, time.perf_counter() - t0)
    ent = {'workload': 'C1 linear BEAR lag 5 on the bundled ysd1 table (1365 rows, 3 groups), bear_lin_bear.cfg: TSV -> pack -> '
                       '%d Adam steps of one 1365-row batch -> heldout evaluation' % epochs,
           'seconds_end_to_end_from_tsv': dt, 'rows_per_s': K * (epochs + 1) / dt, 'heldout_perplexity_bear': perp,
           'roofline': {'bound': 'launch latency (1365 rows per step)', 'frac': None}}
    if cpu:
        from oracle import bear_oracle as O
        ce = 300

        def c1_cpu():
            kms, cnt = O.read_tsv(path, 3)
            oh = O.one_hot(kms)
            gen = torch.Generator().manual_seed(10)
            params = O.init_linear(5, 4, gen)
            hs = torch.zeros((), dtype=torch.float64)
            O.train([(oh, torch.tensor(cnt[:, 0]))] * ce, len(kms), 'linear', params, hs, 0.01, False)
            f = O.ar_linear(oh, params)
            O.evaluation([(oh, f, torch.tensor(cnt[:, 1]), torch.tensor(cnt[:, 0]))], torch.exp(hs), np.array([0.1, 1.0, 10.0]))
        t0 = time.perf_counter()
        c1_cpu()
        cdt = time.perf_counter() - t0
        ent['cpu_baseline'] = {'value': K * (ce + 1) / cdt, 'unit': 'k-mer transition rows/s', 'cores': os.cpu_count() or 1,
                               'kind': 'port', 'sample': 'the same TSV end to end, %d Adam steps instead of %d' % (ce, epochs)}
    out['C1'] = ent
    barrier(c)

    # ---- C2: bear_ref (empirical reference transitions, stop net), lag 10, all 4^10 k-mers, evaluation only ----
    K2 = 1 << 20
    lo, m = shard_range(K2, rank, world)
    k2, c2, s2 = synth_table(c, m, 10, 3, 21, 3, lo)            # dense counts, rows in k-mer order: the 4^10 codes enumerated
    t2 = dl.KmerTable.from_device(k2, c2, m, 10, 'dna')
    d2 = dataset(t2, m, K2, K2, lo)
    _, hs2, af2 = bear_ref._create_params(10, 4, ar_funcs.make_ar_func_stop, {})
    sec = timed(lambda: bear_ref.evaluation(d2, 0, 1, 2, 'dna', torch.exp(hs2), af2, np.array([0.1, 1.0, 10.0]), seed=5), reps=3)
    ent = {'workload': 'C2 bear_ref lag 10, 1 data group + reference column, 4^10 k-mers (dense counts), evaluation only '
                       '(heldout: train + test + reference columns)',
           'rows_per_s': K2 / sec, 'ms': sec * 1e3,
           'roofline': {'bound': 'hbm', 'bytes_per_row': 68, 'frac': 68.0 * K2 / world / sec / 1e9 / peak}}
    if cpu:
        dt, reps = time_cpu(cpu_step_fn(1 << 18, lag=10, dense=True, ref=True), budget_s=5.0)
        ent['cpu_baseline'] = {'value': (1 << 18) / dt, 'unit': 'k-mer transition rows/s', 'cores': os.cpu_count() or 1, 'kind': 'port',
                               'sample': '2^18 dense lag-10 rows x %d evaluation passes (strings -> one-hot -> JC head -> evaluation)' % reps}
    out['C2'] = ent
    del k2, c2, t2, d2

    # ---- C3: linear BEAR, lag 13, 8 groups, ~5e7 k-mers, trained for a fixed number of optimizer steps ----
    K3, B3 = 50_000_000, 1 << 20
    lo, m = shard_range(K3, rank, world)
    k3, c3, s3 = synth_table(c, m, 13, 8, 22, 0, lo)
    t3 = dl.KmerTable.from_device(k3, c3, m, 13, 'dna')
    d3 = dataset(t3, m, B3, K3, lo)
    steps3 = len(d3.ranges)
    sec = timed(lambda: bear_net.train(d3.repeat(2), K3, 2, 3, 'dna', 13, ar_funcs.make_ar_func_linear, {}, 0.01, 'Adam', False))
    ent = {'workload': 'C3 linear BEAR lag 13, 8 groups, 5e7 synthetic k-mers (sparse counts): %d Adam steps of 2^20-row global '
                       'batches on group 3 (two passes)' % (2 * steps3),
           'rows_per_s': 2 * K3 / sec, 'ms': sec * 1e3, 'optimizer_steps': 2 * steps3,
           'roofline': {'bound': 'hbm', 'bytes_per_row': 28, 'frac': 28.0 * 2 * K3 / world / sec / 1e9 / peak}}
    if cpu:
        dt, reps = time_cpu(cpu_step_fn(1 << 20, lag=13), budget_s=5.0)
        ent['cpu_baseline'] = {'value': 2 * (1 << 20) / dt, 'unit': 'k-mer transition rows/s', 'cores': os.cpu_count() or 1, 'kind': 'port',
                               'sample': '2^20 lag-13 rows x %d steps (train pass + eval pass)' % reps}
    out['C3'] = ent
    del k3, c3, t3, d3

    # ---- C4: CNN BEAR, lag 13, 4 groups, ~2e8 k-mers (sampled with replacement): train + eval + posterior pass ----
    K4 = int(os.environ.get('BEAR_BENCH_C4_ROWS', 200_000_000))
    lo, m = shard_range(K4, rank, world)
    k4, c4, s4 = synth_table(c, m, 13, 4, 24, 1, lo)
    t4 = dl.KmerTable.from_device(k4, c4, m, 13, 'dna')
    d4 = dataset(t4, m, 1 << 24, K4, lo)
    kw = {'filter_width': 3}
    state = {}

    def c4_train():
        torch.manual_seed(4)
        state['params'], state['hs'], state['af'] = bear_net.train(d4, K4, 1, 0, 'dna', 13, ar_funcs.make_ar_func_cnn, kw, 0.01,
                                                                   'Adam', False)
    sec_t = timed(c4_train, warm=0)
    sec_e = timed(lambda: bear_net.evaluation(d4, 0, 1, 'dna', torch.exp(state['hs']), state['af'], np.array([0.1, 1.0, 10.0]),
                                              seed=3), warm=0)
    # posterior pass of get_var_probs.get_pdf: 41 Monte-Carlo draws of the transition probabilities of 2^14 query k-mers under
    # 1 BEAR + 3 BMM models (replicas only: every rank scores its own queries)
    Q = 1 << 14
    qk = t4.kmers_str(0, Q)
    qc = torch.stack([c4[g, :, :Q].t() for g in range(4)], 1).to(torch.float64)
    sec_p = timed(lambda: gvp.get_pdf(qk, qc, [float(torch.exp(state['hs']))], state['af'], 41, [0.1, 1.0, 10.0], 0, 'dna', False,
                                      output='numpy', seed=9), warm=0)
    flops = 3 * 2 * 11 * 30 * 16
    ent = {'workload': 'C4 CNN BEAR lag 13 (filter width 3, 30 filters, width 16), 4 groups, %d synthetic k-mers (dense counts): one '
                       'training pass (%d Adam steps of 2^24-row batches) + heldout evaluation + get_pdf posterior pass '
                       '(2^14 query k-mers x 41 draws x 4 models per rank)' % (K4, len(d4.ranges)),
           'train_rows_per_s': K4 / sec_t, 'eval_rows_per_s': K4 / sec_e, 'rows_per_s': 2 * K4 / (sec_t + sec_e),
           'posterior_kmers_per_s': Q * world / sec_p, 'ms': {'train': sec_t * 1e3, 'eval': sec_e * 1e3, 'posterior': sec_p * 1e3},
           'roofline': {'bound': 'fp64 tensor (DMMA) / issue', 'achieved_tflops_dense_layer1': flops * K4 / world / sec_t / 1e12,
                        'hbm_frac_train': 28.0 * K4 / world / sec_t / 1e9 / peak}}
    if cpu:
        dt, reps = time_cpu(cpu_step_fn(1 << 16, lag=13, head='cnn', dense=True), budget_s=5.0)
        ent['cpu_baseline'] = {'value': 2 * (1 << 16) / dt, 'unit': 'k-mer transition rows/s', 'cores': os.cpu_count() or 1, 'kind': 'port',
                               'sample': '2^16 dense lag-13 rows x %d steps (CNN train pass + eval pass)' % reps}
    out['C4'] = ent
    return out


_RESULT_OUT = None               # the process's original stdout, saved by main() before fd 1 is pointed at stderr


def print_result(line):
    out = _RESULT_OUT or sys.stdout
    out.write(line + '\n')
    out.flush()


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
    run_ours(args, make_ctx(rank, world_size, local_rank))


if __name__ == '__main__':
    main()
