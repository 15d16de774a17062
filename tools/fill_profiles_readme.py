"""profiles/README.md = profiles/_tmpl/README.r2.tmpl.md with the numbers of the profiles/r2_* files filled in
+ the round-1 write-up (profiles/_tmpl/README.r1.md)."""
import collections
import csv
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = os.path.join(ROOT, 'profiles')


def sci(x):
    m, e = ('%.3e' % x).split('e')
    sup = str(int(e)).translate(str.maketrans('-0123456789', '⁻⁰¹²³⁴⁵⁶⁷⁸⁹'))
    return '%s·10%s' % (m, sup)


def raw(name):
    rows = list(csv.reader(open(os.path.join(P, name))))
    hdr = rows[0]
    return [dict(zip(hdr, r)) for r in rows[2:]]


def launches():
    rows = list(csv.reader(open(os.path.join(P, 'r2_launches.csv'))))
    hdr, agg = None, collections.defaultdict(float)
    for r in rows:
        if 'Kernel Name' in r:
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            try:
                v = float(r[hdr.index('Metric Value')].replace(',', ''))
            except ValueError:
                continue
            v *= {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}.get(r[hdr.index('Metric Unit')], 1.0)
            agg[r[hdr.index('Kernel Name')]] += v
    tot = sum(agg.values())
    share = lambda key: 100.0 * sum(v for k, v in agg.items() if key in k) / tot
    return share('eval_tile_kernel'), share('linear_train_tc_kernel')


def ktable():
    recs = [json.loads(l) for l in open(os.path.join(P, 'r2_kernel_bench.jsonl')) if l.startswith('{')]
    out = ['| kernel [table] | rows/s | of the HBM roof | note |', '|---|---|---|---|']
    for r in recs:
        frac = r.get('frac_of_measured_hbm_peak')
        out.append('| `%s` | %s | %s | %s |' % (r['kernel'], sci(r['rows_per_s']), ('%.1f %%' % (100 * frac)) if frac and frac > 0.005 else '—',
                                              r.get('note', '')))
    return '\n'.join(out)


def scale(b1):
    out = ['| GPUs | rows/s (train + eval) | ms per step | vs N = 1 | e2e rows/s | e2e ÷ copy-only rate of the measured H2D ceiling | global batch 2²² rows: train rows/s (CUDA graph: kernels + NCCL + optimizer) |',
           '|---|---|---|---|---|---|---|']
    for n, path in ((1, 'r2_bench_full_2p31rows.json'), (2, 'r2_scale_n2.json'), (8, 'r2_scale_n8.json')):
        f = os.path.join(P, path)
        if not os.path.exists(f) or os.path.getsize(f) == 0:
            continue
        d = json.load(open(f))
        bs = {b['global_batch_rows']: b for b in d.get('batch_scaling', [])}
        small = bs.get(1 << 22)
        out.append('| %d | %s | %.1f | %.2f× (%.0f %%) | %s | %.2f (%.0f GB/s aggregate) | %s |' % (
            n, sci(d['value']), d['ms_per_step'], d['value'] / b1['value'], 100 * d['value'] / b1['value'] / n, sci(d['e2e']['value']),
            d['e2e']['h2d_ceiling']['e2e_frac_of_copy_only'], d['e2e']['h2d_ceiling']['aggregate_gbs'],
            ('%s (%.0f µs per optimizer step)' % (sci(small['train_rows_per_s']), small['us_per_optimizer_step'])) if small else '—'))
        pr = d.get('probe')
        if pr and n > 1:
            out.append('| | probe table (2²⁰ rows) recomputed on one rank: loss %.1e, gradient %.1e (rel. to largest), log-likelihoods %.1e, accuracy counts %s | | | | | |' % (
                pr['loss_rel'], pr['grad_rel_to_largest'], pr['ll_rel'], 'equal' if pr['accuracy_counts_equal'] else 'DIFFER'))
    return '\n'.join(out)


def main():
    b = json.load(open(os.path.join(P, 'r2_bench_full_2p31rows.json')))
    rl, e2e = b['roofline'], b['e2e']
    rows = b['config']['rows_total']
    evalms = b['ms_per_step'] - rl['kernel_ms']
    t = raw('r2_train_tc_raw.csv')[0]
    e = raw('r2_eval_tile_raw.csv')[0]
    f = lambda d, k: float(d[k].replace(',', ''))
    nrows = 1 << 26
    stall = lambda d, k: f(d, 'smsp__average_warps_issue_stalled_%s_per_issue_active.ratio' % k)
    le, lt = launches()
    rep = {
        'VALUE': sci(b['value']), 'MS': '%.1f' % b['ms_per_step'], 'TRAINMS': '%.1f' % rl['kernel_ms'], 'GBS': '%.0f' % rl['achieved'],
        'FRAC': '%.1f' % (100 * rl['frac']), 'EVALMS': '%.1f' % evalms, 'EVALRATE': sci(rows / evalms * 1e3), 'E2E': sci(e2e['value']),
        'E2EFRAC': '%.2f' % e2e['h2d_ceiling']['e2e_frac_of_copy_only'], 'H2D': '%.1f' % e2e['h2d_ceiling']['aggregate_gbs'],
        'CPU': sci(b['cpu_baseline']['value']), 'LEVAL': '%.1f' % le, 'LTRAIN': '%.1f' % lt,
        'TSHARE': '%.0f' % (100 * rl['kernel_ms'] / b['ms_per_step']),
        'T_MS': '%.2f' % f(t, 'gpu__time_duration.sum'), 'T_RATE': sci(nrows / f(t, 'gpu__time_duration.sum') * 1e3),
        'T_BPR': '%.2f' % ((f(t, 'dram__bytes_read.sum') * 1e9 + f(t, 'dram__bytes_write.sum') * 1e6) / nrows),
        'T_WF': '%.1f' % (f(t, 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum') / nrows),
        'T_SM': '%.0f' % f(t, 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed'),
        'T_INST': '%.1f' % (f(t, 'smsp__inst_executed.sum') / nrows), 'T_ISSUE': '%.0f' % f(t, 'smsp__issue_active.avg.pct_of_peak_sustained_active'),
        'T_FP64': '%.0f' % f(t, 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active'),
        'T_WAIT': '%.1f' % stall(t, 'wait'), 'T_MPT': '%.1f' % stall(t, 'math_pipe_throttle'), 'T_SSB': '%.1f' % stall(t, 'short_scoreboard'),
        'E_MS': '%.2f' % f(e, 'gpu__time_duration.sum'), 'E_RATE': sci(nrows / f(e, 'gpu__time_duration.sum') * 1e3),
        'E_ISSUE': '%.0f' % f(e, 'smsp__issue_active.avg.pct_of_peak_sustained_active'),
        'E_FP64': '%.0f' % f(e, 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active'),
        'E_INST': '%.1f' % (f(e, 'smsp__inst_executed.sum') / nrows),
        'E_WF': '%.1f' % (f(e, 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum') / nrows),
        'KTABLE': ktable(), 'SCALE': scale(b),
    }
    s = open(os.path.join(P, '_tmpl', 'README.r2.tmpl.md')).read()
    for k, v in rep.items():
        s = s.replace('@%s@' % k, v)
    assert '@' not in s.replace('@', '', 0) or s.count('@') == 0, [w for w in s.split() if w.startswith('@')][:5]
    s += open(os.path.join(P, '_tmpl', 'README.r1.md')).read()
    open(os.path.join(P, 'README.md'), 'w').write(s)
    print('profiles/README.md written')


if __name__ == '__main__':
    main()
