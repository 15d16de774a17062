"""K-mer transition count tables from sequence files, counted on the GPU.

The reference's ``bear_model/summarize.py`` builds the count table with the external KMC binaries
(stages 1-2, summarize.py:103-373) and a Python heap merge (stage 3, summarize.py:380-622).  Its output
is *defined* by brute-force counting over ``'[' * lag + seq + ']'`` (tests/test_summarize.py:96-114);
this module produces that table directly: every transition of every sequence is one GPU thread that
bumps ``counts[kmer][group][next]`` in a device hash table keyed by the packed 2-bit k-mer code
(``bear_count_transitions``).  No KMC, no temporary FASTQ files.

The command line keeps the reference's arguments (``file out_prefix -l -nf -r -mf``; the KMC-specific
``-mk -p -t -pr -s12 -s3`` are accepted and ignored) and output naming
(``<out_prefix>[_rev]_lag_<l>_file_<i>.tsv``, rows ``kmer \\t [[group 0 counts],[group 1 counts],...]``);
``count_kmers`` returns the packed ``KmerTable`` directly, skipping the text round trip.
"""
import argparse
import csv
import json
import os

import numpy as np
import torch

from . import _lib
from ._lib import lib, check, ptr
from .dataloader import KmerTable

alphabet = {'A': 0, 'C': 1, 'G': 2, 'T': 3, ']': 4}      # summarize.py: column order A, C, G, T, $


def read_sequences(path, file_type):
    """Sequences of a FASTA ('fa') or FASTQ ('fq') file (summarize.py:94-100 uses Biopython)."""
    seqs = []
    with open(path) as fh:
        if file_type == 'fa':
            cur = []
            for line in fh:
                line = line.strip()
                if line.startswith('>'):
                    if cur:
                        seqs.append(''.join(cur))
                    cur = []
                elif line:
                    cur.append(line)
            if cur:
                seqs.append(''.join(cur))
        elif file_type == 'fq':
            lines = [l.rstrip('\n') for l in fh]
            i = 0
            while i < len(lines):
                if lines[i].startswith('@') and i + 1 < len(lines):
                    seqs.append(lines[i + 1].strip())
                    i += 4
                else:
                    i += 1
        else:
            raise ValueError("file type must be 'fa' or 'fq'")
    return seqs


def count_kmers(seqs, groups, lag, num_groups=None, reverse=False):
    """Count table of lag-``lag`` transitions.  seqs: list of str (ACGT); groups: int per sequence.
    ``reverse=True`` also counts the reverse complement of every sequence (summarize.py ``-r``).
    Returns (KmerTable resident on the device, stats dict)."""
    dev = _lib.device()
    groups = np.asarray(groups, dtype=np.int32)
    G = int(num_groups if num_groups is not None else (groups.max() + 1 if len(groups) else 1))
    if len(groups) != len(seqs):
        raise ValueError('one group id per sequence is required')
    if len(groups) and (groups.min() < 0 or groups.max() >= G):
        raise ValueError('group ids must lie in [0, %d)' % G)     # (they index the count table on the device)
    lens = np.array([len(s) for s in seqs], dtype=np.int64)
    offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    toff = np.concatenate([[0], np.cumsum(lens + 1)]).astype(np.int64)
    ntrans = int(toff[-1])
    nseq = len(seqs)
    if nseq == 0:
        return KmerTable(np.zeros(4, np.uint64), np.zeros((G, 5, 4), np.uint32), 0, lag, 'dna'), \
            {'distinct': 0, 'skipped': 0, 'transitions': 0}
    text = np.frombuffer(''.join(seqs).encode('ascii'), dtype=np.uint8)
    d_seq = torch.from_numpy(text.copy() if text.size else np.zeros(1, np.uint8)).to(dev)
    d_off = torch.from_numpy(offsets).to(dev)
    d_toff = torch.from_numpy(toff[:-1].copy()).to(dev)
    d_grp = torch.from_numpy(groups).to(dev)
    strands = 2 if reverse else 1
    # distinct k-mers <= min(transitions, 4^lag + start-padded prefixes); keep the table at most half full.
    # Device memory: cap * (8 + 20 G) bytes with cap = the next power of two above twice that bound.
    bound = min(ntrans * strands, sum(4 ** i for i in range(lag + 1)))
    cap = 1 << max(4, int(np.ceil(np.log2(2 * bound + 1))))
    keys = torch.full((cap,), -1, dtype=torch.int64, device=dev)
    counts = torch.zeros((cap, G, 5), dtype=torch.int32, device=dev)
    stats = torch.zeros(3, dtype=torch.int64, device=dev)
    check(lib.bear_count_transitions(ptr(d_seq), ptr(d_off), ptr(d_toff), ptr(d_grp), nseq, ntrans, lag, G, int(reverse),
                                     ptr(keys), ptr(counts), cap, ptr(stats), _lib.stream()))
    distinct, skipped, overflow = (int(x) for x in stats.cpu())
    if overflow:
        raise OverflowError('a transition count exceeded 2^32 - 1')
    rows = torch.nonzero(keys != -1).reshape(-1)              # occupied slots
    n = int(rows.numel())
    assert n == distinct
    # The slot a key lands in depends on which thread wins a probe race, so slot order is not reproducible.  Rows go
    # out ordered by a hash of the key instead: the same table on every run, and shuffled as the reference recommends
    # for training (a table sorted by k-mer makes consecutive minibatches share their leading positions).
    mixed = keys[rows] * -7046029254386353131                # 0x9E3779B97F4A7C15 (wraps)
    mixed = mixed ^ (mixed >> 29)
    rows = rows[torch.argsort(mixed * -4658895280553007687)]
    stride = max((n + 3) // 4 * 4, 4)
    out_k = torch.zeros(stride, dtype=torch.int64, device=dev)
    out_c = torch.zeros((G, 5, stride), dtype=torch.int32, device=dev)
    check(lib.bear_gather_table(ptr(keys), ptr(counts), ptr(rows), n, G, stride, ptr(out_k), ptr(out_c), _lib.stream()))
    table = KmerTable.from_device(out_k, out_c, n, lag, 'dna')
    return table, {'distinct': distinct, 'skipped': skipped, 'transitions': ntrans * strands}


def write_tsv(table, out_prefix, lag, max_file_gb=0.1):
    """``<out_prefix>_lag_<lag>_file_<i>.tsv`` in the reference's row format; returns the number of files."""
    k, c = table.device_tensors()
    n = table.num_rows
    kmers = [s.decode() for s in table.kmers_str()]
    counts = c[:, :, :n].permute(2, 0, 1).cpu().numpy()
    limit = max(int(max_file_gb * 1e9), 1)
    fi, written, fh = 0, 0, None
    for i in range(n):
        line = kmers[i] + '\t' + json.dumps(counts[i].tolist(), separators=(',', ':')) + '\n'
        if fh is None or written + len(line) > limit:
            if fh is not None:
                fh.close()
                fi += 1
            fh = open('{}_lag_{}_file_{}.tsv'.format(out_prefix, lag, fi), 'w')
            written = 0
        fh.write(line)
        written += len(line)
    if fh is None:
        fh = open('{}_lag_{}_file_{}.tsv'.format(out_prefix, lag, 0), 'w')
    fh.close()
    return fi + 1


def run(args):
    seqs, groups = [], []
    with open(args.file, newline='') as fh:
        for row in csv.reader(fh):
            if not row:
                continue
            path, group, ftype = row[0].strip(), int(row[1]), row[2].strip()
            s = read_sequences(path, ftype)
            seqs += s
            groups += [group] * len(s)
    n_groups = max(groups) + 1 if groups else 1
    n_bins = 0
    for lag in range(1, args.l + 1):
        table, _ = count_kmers(seqs, groups, lag, num_groups=n_groups, reverse=args.r)
        n_bins = max(n_bins, write_tsv(table, args.out_prefix, lag, args.mf))
    return n_bins


def main(args):
    """Same contract as the reference's ``summarize.main`` (summarize.py:648-665): forward counts unless
    ``-nf``; with ``-r`` a second set ``<out_prefix>_rev_*`` that also counts reverse complements.
    Returns (n_bins, n_bins_rev)."""
    store_r = args.r
    args.r = False
    n_bins = None if args.nf else run(args)
    n_bins_rev = None
    if store_r:
        args.r = True
        args.out_prefix += '_rev'
        n_bins_rev = run(args)
    return n_bins, n_bins_rev


def make_parser():
    parser = argparse.ArgumentParser(description="Preprocess for collapsed BEAR training (GPU k-mer transition counting).")
    parser.add_argument('file', help='Input file: csv of individual files, their group number and type (fa / fq).')
    parser.add_argument('out_prefix', help='Prefix for output files.')
    parser.add_argument('-l', default=10, type=int, help='Maximum lag of BEAR model.')
    parser.add_argument('-mf', default=0.1, type=float, help='Maximum size of final dataset chunks (Gb).')
    parser.add_argument('-nf', action='store_true', default=False, help='Do not compute the forward direction.')
    parser.add_argument('-r', action='store_true', default=False, help='Compute reverse direction.')
    for flag, kw in (('-mk', dict(default=12, type=float)), ('-p', dict(default='')), ('-t', dict(default='tmp/')),
                     ('-pr', dict(action='store_true', default=False)), ('-s12', dict(action='store_true', default=False)),
                     ('-s3', dict(action='store_true', default=False))):
        parser.add_argument(flag, help='accepted for compatibility with the KMC-based reference; ignored', **kw)
    return parser


if __name__ == '__main__':
    main(make_parser().parse_args())
