"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export per CUDA source line:
share of warp-stall samples and of executed instructions.  usage: ncu_lines.py file.csv [top]"""
import csv
import sys


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    rows = list(csv.reader(open(path)))
    items, cur = [], None
    for r in rows:
        if len(r) >= 2 and r[0] == 'File Path':
            cur = r[1].split('/')[-1]
            continue
        if len(r) < 8 or r[0] in ('', 'Line No', 'Function Name'):
            continue
        try:
            inst, samp = int(r[7]), int(r[6])
        except ValueError:
            continue
        items.append((samp, inst, cur, r[0], r[1].strip()[:100]))
    ts, ti = sum(i[0] for i in items), sum(i[1] for i in items)
    print('samples %d, warp instructions %d' % (ts, ti))
    for it in sorted(items, reverse=True)[:top]:
        print('%5.1f%% samp %5.1f%% inst  %s:%s  %s' % (100 * it[0] / ts, 100 * it[1] / ti, it[2], it[3], it[4]))


if __name__ == '__main__':
    main()
