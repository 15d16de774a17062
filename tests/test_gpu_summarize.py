"""GPU k-mer transition counting (bear_b200.summarize) against the brute-force definition of the count
table that the reference's own test uses (bear_model/tests/test_summarize.py:96-114), on the reference's
example FASTA / FASTQ files (tests/golden/exdata, copied from bear_model/tests/exdata) and on random reads."""
import csv
import json
import os
from collections import defaultdict

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu

EXDATA = os.path.join(GOLDEN, 'exdata')
COMP = {'A': 'T', 'C': 'G', 'G': 'C', 'T': 'A'}


def brute_force(seqs, groups, n_groups, lag, reverse):
    """tests/test_summarize.py:96-114"""
    alphabet = {'A': 0, 'C': 1, 'G': 2, 'T': 3, ']': 4}
    counts = defaultdict(lambda: [[0] * 5 for _ in range(n_groups)])
    for seq, g in zip(seqs, groups):
        variants = [seq] + ([''.join(COMP[c] for c in reversed(seq))] if reverse else [])
        for s in variants:
            full = '[' * lag + s + ']'
            for j in range(lag, len(full)):
                counts[full[j - lag:j]][g][alphabet[full[j]]] += 1
    return dict(counts)


def table_to_dict(table):
    k, c = table.device_tensors()
    n = table.num_rows
    kmers = [s.decode() for s in table.kmers_str()]
    counts = c[:, :, :n].permute(2, 0, 1).cpu().numpy()
    assert len(set(kmers)) == len(kmers)
    return {km: counts[i].tolist() for i, km in enumerate(kmers)}


def test_reference_example_files_all_lags(cuda, tmp_path):
    """summarize.main on the reference's example inputs: forward and '-r' outputs for lags 1..10."""
    from bear_b200 import summarize
    files = [('infile_0.fa', 0, 'fa'), ('infile_1.fq', 0, 'fq'), ('infile_2.fq', 2, 'fq'), ('infile_3.fa', 1, 'fa'),
             ('infile_4.fq', 1, 'fq')]
    listing = tmp_path / 'infiles.csv'
    with open(listing, 'w') as fh:
        for name, g, t in files:
            fh.write('{},{},{}\n'.format(os.path.join(EXDATA, name), g, t))
    seqs, groups = [], []
    for name, g, t in files:
        s = summarize.read_sequences(os.path.join(EXDATA, name), t)
        seqs += s
        groups += [g] * len(s)
    assert [len(summarize.read_sequences(os.path.join(EXDATA, n), t)) for n, _, t in files] == [3, 2, 2, 4, 2]
    args = summarize.make_parser().parse_args([str(listing), str(tmp_path / 'out'), '-l', '10', '-r', '-mf', '2'])
    nbins, nbins_rev = summarize.main(args)
    assert nbins == 1 and nbins_rev == 1
    for rev, prefix in ((False, str(tmp_path / 'out')), (True, str(tmp_path / 'out') + '_rev')):
        for lag in range(1, 11):
            want = brute_force(seqs, groups, 3, lag, rev)
            got = {}
            with open('{}_lag_{}_file_0.tsv'.format(prefix, lag), newline='') as fh:
                for kmer, cs in csv.reader(fh, delimiter='\t'):
                    assert kmer not in got
                    got[kmer] = json.loads(cs)
            assert got == want


@pytest.mark.parametrize('lag', [1, 5, 13, 20])
def test_random_reads_match_brute_force_and_feed_the_dataloader(cuda, lag, tmp_path):
    from bear_b200 import dataloader, summarize
    rng = np.random.default_rng(lag)
    seqs = [''.join(rng.choice(list('ACGT'), size=int(rng.integers(1, 60)))) for _ in range(400)]
    seqs[7] = seqs[7][:3] + 'N' + seqs[7][3:]               # a read with an ambiguous base
    groups = rng.integers(0, 3, size=len(seqs)).tolist()
    table, stats = summarize.count_kmers(seqs, groups, lag, num_groups=3, reverse=True)
    got = table_to_dict(table)
    clean = [s for s in seqs if 'N' not in s]
    cgroups = [g for s, g in zip(seqs, groups) if 'N' not in s]
    want = brute_force(clean, cgroups, 3, lag, True)
    # transitions touching the ambiguous base are skipped; everything else of that read still counts
    extra = brute_force([seqs[7].replace('N', 'A')], [groups[7]], 3, lag, True)
    assert stats['skipped'] > 0 and stats['distinct'] == len(got)
    for kmer, c in want.items():
        base = np.array(c)
        assert kmer in got
        assert np.all(np.array(got[kmer]) >= base) and np.all(np.array(got[kmer]) <= base + np.array(extra.get(kmer, base * 0)))
    assert sum(np.sum(c) for c in got.values()) == stats['transitions'] - stats['skipped']
    # the TSV it writes is what the packer reads back
    summarize.write_tsv(table, str(tmp_path / 'o'), lag, 2)
    back = dataloader.KmerTable.from_file(str(tmp_path / 'o') + '_lag_%d_file_0.tsv' % lag, 'dna', 3)
    assert table_to_dict(back) == got
