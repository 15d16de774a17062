"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export per kernel and CUDA source line:
share of warp-stall samples, of executed instructions and shared-memory wavefronts.
usage: ncu_lines.py file.csv [top]"""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    kernels = collections.OrderedDict()
    kern, cur, hdr = None, None, None
    for r in csv.reader(open(path)):
        if len(r) >= 2 and r[0] in ('Function Name', 'Kernel Name'):
            kern = re.sub(r'\(.*', '', r[1]).replace('void ', '').replace('<unnamed>::', '')
            kernels.setdefault(kern, collections.defaultdict(lambda: [0, 0, 0.0, 0.0, '']))
            continue
        if len(r) >= 2 and r[0] == 'File Path':
            cur = r[1].split('/')[-1]
            continue
        if len(r) > 3 and r[0] == 'Line No':
            hdr = r
            continue
        if kern is None or hdr is None or len(r) < 8 or r[0] == '':
            continue
        try:
            inst, samp, ln = int(r[7]), int(r[6]), int(r[0])
        except ValueError:
            continue
        w = wi = 0.0
        if 'L1 Wavefronts Shared' in hdr and len(r) >= len(hdr):
            try:
                w = float(r[hdr.index('L1 Wavefronts Shared')] or 0)
                wi = float(r[hdr.index('L1 Wavefronts Shared Ideal')] or 0)
            except ValueError:
                pass
        e = kernels[kern][(cur, ln)]
        e[0] += samp
        e[1] += inst
        e[2] += w
        e[3] += wi
        e[4] = r[1].strip()[:90]
    for kern, lines in kernels.items():
        ts = sum(v[0] for v in lines.values()) or 1
        ti = sum(v[1] for v in lines.values()) or 1
        tw = sum(v[2] for v in lines.values())
        twi = sum(v[3] for v in lines.values())
        print('== %s: %d samples, %d warp instructions, %.0f shared wavefronts (ideal %.0f)' % (kern, ts, ti, tw, twi))
        for (f, ln), v in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
            print('%5.1f%% samp %5.1f%% inst %5.1f%% smem-wf  %s:%d  %s' % (100 * v[0] / ts, 100 * v[1] / ti,
                                                                           100 * v[2] / tw if tw else 0.0, f, ln, v[4]))
        print()


if __name__ == '__main__':
    main()
