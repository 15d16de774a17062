"""CPU tests of the boundary: the C-ABI library loads and exports every symbol include/bear_b200.h
declares (no compute calls without a GPU), and the host packer reproduces the reference's text
formats bit-exactly (dataloader.py:6-109, core.py:142-174)."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, SPARSE, YSD1
from oracle import bear_oracle as O


def test_library_exports_every_declared_symbol():
    from bear_b200 import _lib
    header = open(os.path.join(ROOT, 'include', 'bear_b200.h')).read()
    header = re.sub(r'/\*.*?\*/', '', header, flags=re.S)
    declared = sorted(set(re.findall(r'\b(bear_[a-z0-9_]+)\s*\(', header)))
    assert len(declared) >= 28
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(raw, name), 'symbol %s declared in bear_b200.h is not exported' % name
    assert set(_lib.EXPORTS) == set(declared), 'ctypes signatures and header disagree'
    assert _lib.lib.bear_version() == 1
    assert [_lib.lib.bear_alphabet_size(a) for a in (0, 1, 2, 3)] == [4, 4, 20, -1]
    assert [_lib.lib.bear_max_lag(a) for a in (0, 1, 2)] == [29, 29, 12]


def test_missing_library_fails_loudly(tmp_path, monkeypatch):
    """No fallback: a missing .so is an ImportError with a build hint."""
    import importlib
    from bear_b200 import _lib
    monkeypatch.setattr(_lib, 'LIB_PATH', str(tmp_path / 'nope.so'))
    with pytest.raises(ImportError, match='no CPU or pure-PyTorch fallback'):
        _lib._load()


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, 'bear_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.cpp', '.h')):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle', src, flags=re.M), f


def test_compute_entry_points_need_cuda():
    import torch
    from bear_b200 import _lib
    if torch.cuda.is_available():
        pytest.skip('CPU-only check')
    with pytest.raises(_lib.BearError, match='no CPU fallback'):
        _lib.device()


# ------------------------------------------------------------------------------------------------
def test_pack_tsv_matches_oracle_reader_bit_exact():
    from bear_b200 import dataloader as dl
    t = dl.KmerTable.from_file(YSD1, 'dna', 3)
    kmers, counts = O.read_tsv(YSD1, 3)
    assert t.num_rows == 1365 == dl.count_rows(YSD1) and t.lag == 5 and t.stride % 4 == 0
    assert [k.decode() for k in t.kmers_str()] == kmers
    got = np.transpose(t.counts_host[:, :, :t.num_rows], (2, 0, 1))
    assert np.array_equal(got.astype(np.float64), counts)
    assert not t.counts_host[:, :, t.num_rows:].any()          # padding rows are zero


def test_pack_sparse_matches_oracle_reader():
    from bear_b200 import dataloader as dl
    t = dl.KmerTable.from_file(SPARSE, 'dna', 1, sparse=True)
    kmers, counts = O.read_sparse(SPARSE, 1)
    assert [k.decode() for k in t.kmers_str()] == kmers
    assert np.array_equal(np.transpose(t.counts_host[:, :, :t.num_rows], (2, 0, 1)).astype(float), counts)
    # the dense TSV twins of the same toy sequences agree row for row
    twin = dl.KmerTable.from_file(os.path.join(os.path.dirname(SPARSE), 'kmaps', 'ex_seqs_lag_3_file_0.tsv'), 'dna', 1)
    a = dict(zip(t.kmers_str().tolist(), t.counts_host[0, :, :t.num_rows].T.tolist()))
    b = dict(zip(twin.kmers_str().tolist(), twin.counts_host[0, :, :twin.num_rows].T.tolist()))
    assert a == b


@pytest.mark.parametrize('alphabet,lag', [('dna', 1), ('dna', 13), ('dna', 20), ('dna', 29), ('rna', 7), ('prot', 12), ('prot', 3)])
def test_encode_decode_round_trip(alphabet, lag):
    from bear_b200 import dataloader as dl
    rng = np.random.default_rng(lag)
    letters = O.ALPHABETS_IN[alphabet][:-1]
    kmers = []
    for _ in range(500):
        ns = int(rng.integers(0, lag + 1)) if rng.random() < 0.3 else 0
        kmers.append('[' * ns + ''.join(rng.choice(letters, size=lag - ns)))
    codes, got_lag = dl.encode_kmers(kmers, alphabet)
    assert got_lag == lag
    assert [k.decode() for k in dl.decode_kmers(codes, lag, alphabet)] == kmers
    if alphabet != 'prot':
        # numeric order of the payload = lexicographic order of start-free k-mers
        plain = sorted(k for k in kmers if '[' not in k)
        order = {'A': 0, 'C': 1, 'G': 2, 'T': 3, 'U': 3}
        pc, _ = dl.encode_kmers(plain, alphabet) if plain else (np.zeros(0, np.uint64), lag)
        assert np.all(np.diff(pc.astype(np.int64)) >= 0)
        for k, c in zip(plain[:20], pc[:20]):
            assert int(c) == sum(order[ch] << (2 * (lag - 1 - j)) for j, ch in enumerate(k))


def _write(tmp_path, text, name='t.tsv'):
    p = tmp_path / name
    p.write_text(text)
    return str(p)


def test_pack_error_cases(tmp_path):
    from bear_b200 import dataloader as dl
    from bear_b200._lib import BearError
    ok = 'ACG\t[[1,2,3,4,5]]\n'
    cases = {
        'outside the alphabet': 'ANG\t[[1,2,3,4,5]]\n',
        "start symbol '\\[' after a letter": 'A[G\t[[1,2,3,4,5]]\n',
        'length 2 differs from 3': ok + 'AC\t[[1,2,3,4,5]]\n',
        'negative or not an integer': 'ACG\t[[1,2.5,3,4,5]]\n',
        'exceeds uint32': 'ACG\t[[1,2,3,4,5000000000]]\n',
        'no tab separator': 'ACG [[1,2,3,4,5]]\n',
        "expected ','": 'ACG\t[[1,2,3,4]]\n',
        'wrong alphabet size': 'ACG\t[[1,2,3,4,5,6]]\n',
        'wrong num_ds': 'ACG\t[[1,2,3,4,5],[1,2,3,4,5]]\n',
    }
    for msg, text in cases.items():
        with pytest.raises(BearError, match=msg):
            dl.KmerTable.from_file(_write(tmp_path, text), 'dna', 1)
    with pytest.raises(BearError, match='cannot open'):
        dl.KmerTable.from_file(str(tmp_path / 'absent.tsv'), 'dna', 1)
    with pytest.raises(BearError, match='exceeds the packed layout'):
        dl.KmerTable.from_file(_write(tmp_path, 'A' * 30 + '\t[[1,2,3,4,5]]\n'), 'dna', 1)


def test_pack_edge_inputs(tmp_path):
    from bear_b200 import dataloader as dl
    empty = dl.KmerTable.from_file(_write(tmp_path, ''), 'dna', 2)
    assert empty.num_rows == 0 and len(dl.KmerDataset(empty, 10)) == 0
    # header line, blank lines, CRLF, float-formatted integers, maximum count
    text = 'kmer\tcounts\n\nACGT\t[[1.0, 2, 3e0, 4294967295, 0],[0,0,0,0,0]]\r\n[[[A\t[[0,0,0,0,0],[7,0,0,0,1]]\n'
    t = dl.KmerTable.from_file(_write(tmp_path, text), 'dna', 2, header=True)
    assert t.num_rows == 2 and t.lag == 4
    assert t.counts_host[0, :, 0].tolist() == [1, 2, 3, 4294967295, 0]
    assert t.counts_host[1, :, 1].tolist() == [7, 0, 0, 0, 1]
    assert [k.decode() for k in t.kmers_str()] == ['ACGT', '[[[A']
    assert int(t.kmers_host[1] >> np.uint64(58)) == 3
    ds = dl.KmerDataset(t, 1).repeat(3)
    assert len(ds) == 6 and [b[:2] for b in ds.batches()] == [(0, 1), (1, 1)] * 3
    # integer-valued decimals (json.dumps of float arrays) take the packer's fast path; a fractional part is an error
    text = 'ACGT\t[[5.000, 7., 12.0, 4294967295.0, 1.0e1]]\nTTTT\t[[0.0, 0.00 , 3.0,1.0 ,2]]\n'
    t = dl.KmerTable.from_file(_write(tmp_path, text, 'dec.tsv'), 'dna', 1)
    assert t.counts_host[0, :, 0].tolist() == [5, 7, 12, 4294967295, 10] and t.counts_host[0, :, 1].tolist() == [0, 0, 3, 1, 2]
    for bad in ('ACGT\t[[1.05,0,0,0,0]]\n', 'ACGT\t[[1.0.0,0,0,0,0]]\n', 'ACGT\t[[4294967296.0,0,0,0,0]]\n'):
        with pytest.raises(Exception):
            dl.KmerTable.from_file(_write(tmp_path, bad, 'bad_dec.tsv'), 'dna', 1)


def test_invalid_symbol_policy(tmp_path):
    """DNA k-mers with a symbol outside the alphabet: an error naming the k-mer by default, dropped (and counted) with
    on_invalid='skip'; the other rows are unchanged.  The policy does not leak into later loads."""
    from bear_b200 import _lib, dataloader as dl
    rows = ['ACGTA\t[[1,2,3,4,5]]', 'ACNTA\t[[9,9,9,9,9]]', '[[CGT\t[[0,0,7,0,1]]', 'NNNNN\t[[1,1,1,1,1]]', 'TTTTT\t[[5,4,3,2,1]]']
    path = _write(tmp_path, '\n'.join(rows) + '\n', 'n.tsv')
    with pytest.raises(_lib.BearError, match='outside the alphabet'):
        dl.KmerTable.from_file(path, 'dna', 1)
    t = dl.KmerTable.from_file(path, 'dna', 1, on_invalid='skip')
    assert t.num_rows == 3 and t.skipped_rows == 2
    assert [k.decode() for k in t.kmers_str()] == ['ACGTA', '[[CGT', 'TTTTT']
    assert t.counts_host[0, :, :3].T.tolist() == [[1, 2, 3, 4, 5], [0, 0, 7, 0, 1], [5, 4, 3, 2, 1]]
    with pytest.raises(_lib.BearError, match='outside the alphabet'):
        dl.KmerTable.from_file(path, 'dna', 1)
    clean = dl.KmerTable.from_file(YSD1, 'dna', 3, on_invalid='skip')
    assert clean.skipped_rows == 0 and clean.num_rows == 1365


def test_dataset_batching_and_sharding_cover_every_row_once():
    from bear_b200 import dataloader as dl
    t = dl.KmerTable.from_file(YSD1, 'dna', 3)
    ds = dl.KmerDataset(t, 300)
    assert [n for _, n, _ in ds.batches()] == [300, 300, 300, 300, 165]
    for world in (2, 3, 8):
        seen = []
        for rank in range(world):
            sh = ds.shard(rank, world)
            assert [g for _, _, g in sh.batches()] == [300, 300, 300, 300, 165]
            seen.append(sh.table.kmers_str().tolist())
            # per batch the local slices of all ranks add up to the global batch
        assert sorted(sum(seen, [])) == sorted(t.kmers_str().tolist())
        per_batch = np.sum([[n for _, n, _ in ds.shard(r, world).batches()] for r in range(world)], 0)
        assert per_batch.tolist() == [300, 300, 300, 300, 165]


@pytest.mark.parametrize('path,sparse,num_ds', [(YSD1, False, 3), (SPARSE, True, 1)])
def test_rank_sharded_ingest_equals_shard_of_the_full_table(path, sparse, num_ds):
    """bear_pack_shard: every rank parses only its slice of every global batch; the result is bit-identical to
    KmerDataset.shard() of the whole table (same rows, same order, same batch geometry and global row ids)."""
    from bear_b200 import dataloader as dl
    full = dl.KmerTable.from_file(path, 'dna', num_ds, sparse=sparse)
    K = full.num_rows
    for batch, world in ((455, 3), (100, 4), (max(K, 1), 2), (7, 5), (K + 10, 3)):
        want_ds = dl.KmerDataset(full, batch)
        total = 0
        for rank in range(world):
            t, k_file = dl.KmerTable.from_file_shard(path, 'dna', num_ds, batch, rank, world, sparse=sparse)
            assert k_file == K and t.lag == full.lag
            want = want_ds.shard(rank, world) if world > 1 else want_ds
            assert t.num_rows == want.table.num_rows
            assert np.array_equal(t.kmers_host[:t.num_rows], want.table.kmers_host[:t.num_rows])
            assert np.array_equal(t.counts_host[:, :, :t.num_rows], want.table.counts_host[:, :, :t.num_rows])
            ranges, grows, ids = dl.shard_ranges(K, batch, rank, world)
            assert ranges == want.ranges and grows == want.global_rows and ids == want.row_ids
            total += t.num_rows
        assert total == K


def test_rank_sharded_ingest_of_a_multi_file_dataset(tmp_path):
    """Several files of one dataset (models/train_bear_net.py:78-86): global batches run over the concatenated rows, file
    boundaries fall anywhere inside them; every rank's per-file shards, concatenated, equal shard() of the whole table."""
    from bear_b200 import dataloader as dl
    lines = open(YSD1).read().splitlines()
    cuts = [0, 333, 334, 1000, len(lines)]                       # (a one-row file among them)
    files = []
    for i in range(len(cuts) - 1):
        files.append(_write(tmp_path, '\n'.join(lines[cuts[i]:cuts[i + 1]]) + '\n', 'part_%d.tsv' % i))
    full = dl.KmerTable.from_file(YSD1, 'dna', 3)
    K = full.num_rows
    for batch, world in ((455, 3), (100, 4), (K, 2), (7, 5), (2000, 3)):
        want_ds = dl.KmerDataset(full, batch)
        for rank in range(world):
            parts, off = [], 0
            for f, n in zip(files, np.diff(cuts)):
                t, k_file = dl.KmerTable.from_file_shard(f, 'dna', 3, batch, rank, world, row_offset=off, dataset_rows=K)
                assert k_file == n
                parts.append(t)
                off += int(n)
            got = dl.KmerTable.concat(parts)
            want = want_ds.shard(rank, world).table
            assert got.num_rows == want.num_rows
            assert np.array_equal(got.kmers_host[:got.num_rows], want.kmers_host[:got.num_rows])
            assert np.array_equal(got.counts_host[:, :, :got.num_rows], want.counts_host[:, :, :got.num_rows])
            assert got.num_rows == dl.rank_rows_before(K, K, batch, rank, world)
    with pytest.raises(Exception):                               # a file that does not fit the stated dataset
        dl.KmerTable.from_file_shard(files[-1], 'dna', 3, 100, 0, 2, row_offset=1300, dataset_rows=1365)


def test_packed_binary_cache_round_trip(tmp_path):
    """TSV -> BEARPACK shard -> memory-mapped table: bit-identical arrays and metadata."""
    from bear_b200 import dataloader as dl
    out = str(tmp_path / 'ysd1.bearpack')
    t = dl.pack_files([YSD1], out, 'dna', 3)
    u = dl.KmerTable.load(out)
    assert (u.num_rows, u.lag, u.alphabet, u.num_ds, u.A1, u.stride) == (t.num_rows, t.lag, t.alphabet, t.num_ds, t.A1, t.stride)
    assert np.array_equal(np.asarray(u.kmers_host), t.kmers_host) and np.array_equal(np.asarray(u.counts_host), t.counts_host)
    ds = dl.KmerDataset(u, 500)
    assert [n for _, n, _ in ds.batches()] == [500, 500, 365]
    with pytest.raises(ValueError, match='not a BEARPACK'):
        dl.KmerTable.load(YSD1)


def test_multithreaded_pack_equals_single_thread(tmp_path, monkeypatch):
    """The two-pass multi-threaded parse places every row exactly where the single-threaded one does,
    honours first_row / max_rows windows and reports the earliest error."""
    import ctypes
    from bear_b200 import _lib, dataloader as dl
    rng = np.random.default_rng(0)
    lag, K = 11, 60000
    lines = []
    for i in range(K):
        k = ''.join(rng.choice(list('ACGT'), size=lag))
        c = rng.poisson(0.7, size=(2, 5))
        lines.append(k + '\t[[' + ','.join(map(str, c[0])) + '],[' + ','.join(map(str, c[1])) + ']]')
    path = _write(tmp_path, '\n'.join(lines) + '\n', 'big.tsv')
    assert os.path.getsize(path) > (1 << 20)
    monkeypatch.setenv('BEAR_PACK_THREADS', '1')
    a = dl.KmerTable.from_file(path, 'dna', 2)
    monkeypatch.setenv('BEAR_PACK_THREADS', '7')
    b = dl.KmerTable.from_file(path, 'dna', 2)
    assert a.num_rows == b.num_rows == K == dl.count_rows(path)
    assert np.array_equal(a.kmers_host, b.kmers_host) and np.array_equal(a.counts_host, b.counts_host)
    # a window of rows
    kmers = np.zeros(1000, np.uint64)
    counts = np.zeros((2, 5, 1000), np.uint32)
    rows, lg = ctypes.c_int64(), ctypes.c_int()
    _lib.check(_lib.lib.bear_pack_tsv(path.encode(), 0, 0, 2, 12345, 1000, _lib.ptr(kmers), _lib.ptr(counts), 1000,
                                      ctypes.byref(rows), ctypes.byref(lg)))
    assert rows.value == 1000 and lg.value == lag
    assert np.array_equal(kmers, a.kmers_host[12345:13345]) and np.array_equal(counts, a.counts_host[:, :, 12345:13345])
    # two malformed rows: the earlier one is reported
    lines[40000] = lines[40000].replace('\t', ' ')
    lines[20000] = 'ACGTNACGTAC' + lines[20000][lag:]
    bad = _write(tmp_path, '\n'.join(lines) + '\n', 'bad.tsv')
    with pytest.raises(_lib.BearError, match='outside the alphabet'):
        dl.KmerTable.from_file(bad, 'dna', 2)


def _rank_vectors(A1, nmax):
    """All count vectors of A1 entries with sum <= nmax in the order of the 12-bit rank coding (include/bear_b200.h):
    by sum, lexicographic within one sum."""
    def comps(m, parts):
        if parts == 1:
            return [[m]]
        return [[v] + rest for v in range(m + 1) for rest in comps(m - v, parts - 1)]
    out = [c for m in range(nmax + 1) for c in comps(m, A1)]
    return np.array(out, dtype=np.uint32)


def _decode_compact(buf, esc, n, lag, alphabet, G, A1, wire=8):
    """numpy reader of the compact transfer format (include/bear_b200.h)."""
    bits, start_esc = wire & 15, bool(wire & 16)
    kbits = 5 * lag if alphabet == 'prot' else 2 * lag + (0 if start_esc else 6)
    kb, pitch = (kbits + 7) // 8, (n + 15) // 16 * 16
    b = buf.numpy()
    v = np.zeros(n, dtype=np.uint64)
    for i in range(kb):
        v |= b[i * pitch:i * pitch + n].astype(np.uint64) << np.uint64(8 * i)
    if alphabet != 'prot':
        pay = v & np.uint64((1 << (2 * lag)) - 1)
        v = pay | ((v >> np.uint64(2 * lag)) << np.uint64(58))
    if bits == 12:                      # 12-bit rank of the (row, group) count vector: byte plane + nibble plane per group
        nmax = 10 if A1 == 5 else 3
        vecs = _rank_vectors(A1, nmax)
        counts = np.zeros((G * A1, n), dtype=np.uint32)
        escaped = np.zeros((G, n), dtype=bool)
        for g in range(G):
            base = kb * pitch + g * (pitch + pitch // 2)
            lo8 = b[base:base + n].astype(np.uint32)
            nib = b[base + pitch:base + pitch + pitch // 2].astype(np.uint32)
            hi4 = np.stack([nib & 15, nib >> 4], axis=1).reshape(-1)[:n]
            r = lo8 | (hi4 << 8)
            escaped[g] = r == 4095
            assert np.all((r < len(vecs)) | escaped[g])
            counts[g * A1:(g + 1) * A1, ~escaped[g]] = vecs[r[~escaped[g]]].T
        for pl, row, val in esc.numpy().view(np.uint32).reshape(-1, 3):
            if pl == 0xffffffff:
                assert start_esc and v[row] >> np.uint64(58) == 0
                v[row] |= np.uint64(val) << np.uint64(58)
                continue
            assert escaped[pl // A1, row] and counts[pl, row] == 0 and val != 0
            counts[pl, row] = val
        return v, counts.reshape(G, A1, n)
    if bits == 8:
        counts = np.stack([b[(kb + pl) * pitch:(kb + pl) * pitch + n].astype(np.uint32) for pl in range(G * A1)])
    else:                               # two rows per byte, low nibble = even row
        cp = pitch // 2
        planes = [b[kb * pitch + pl * cp:kb * pitch + (pl + 1) * cp].astype(np.uint32) for pl in range(G * A1)]
        counts = np.stack([np.stack([p & 15, p >> 4], axis=1).reshape(-1)[:n] for p in planes])
    for pl, row, val in esc.numpy().view(np.uint32).reshape(-1, 3):
        if pl == 0xffffffff:            # n_start of a start-padded row
            assert start_esc and v[row] >> np.uint64(58) == 0
            v[row] |= np.uint64(val) << np.uint64(58)
            continue
        assert counts[pl, row] == (255 if bits == 8 else 15)
        counts[pl, row] = val
    return v, counts.reshape(G, A1, n)


@pytest.mark.parametrize('alphabet,lag,n', [('dna', 20, 1000), ('dna', 29, 77), ('dna', 1, 16), ('prot', 12, 333), ('rna', 5, 70000)])
def test_compact_transfer_format_is_lossless(alphabet, lag, n, monkeypatch):
    """Host side of the compact transfer format: k-mer byte planes, 8- or 4-bit count planes and escapes decode back to
    the packed table bit for bit (counts on both sides of 15 and 255, up to 2^32 - 1; start-padded k-mers;
    sub-ranges; threaded and single-threaded)."""
    from bear_b200 import dataloader as dl
    rng = np.random.default_rng(lag * 1000 + n)
    A1 = (20 if alphabet == 'prot' else 4) + 1
    G = 2
    if alphabet == 'prot':
        codes = rng.integers(0, 1 << (5 * lag), size=n, dtype=np.uint64)
    else:
        codes = rng.integers(0, 4 ** lag, size=n, dtype=np.uint64)
        ns = np.where(rng.random(n) < 0.3, rng.integers(1, lag + 1, size=n), 0).astype(np.uint64)
        for i in np.flatnonzero(ns):
            codes[i] &= np.uint64((1 << (2 * (lag - int(ns[i])))) - 1)
        codes |= ns << np.uint64(58)
    counts = rng.integers(0, 4, size=(n, G, A1)) * rng.choice([1, 80, 85, 127, 128, 1000, 1431655765], size=(n, G, A1))
    sparse_rows = rng.random(n) < 0.6              # rows the 12-bit rank coding can hold (small sums) next to rows it escapes
    counts[sparse_rows] = (rng.random((int(sparse_rows.sum()), G, A1)) < 0.25) * rng.integers(1, 4, size=(int(sparse_rows.sum()), G, A1))
    table = dl.KmerTable.from_arrays((codes, lag), counts, alphabet)
    for threads in ('1', '4'):
        monkeypatch.setenv('BEAR_PACK_THREADS', threads)
        for r0, m in ((0, n), (3, n - 7), (n // 2, 1)):
            sub, ksub = table.counts_host[:, :, r0:r0 + m], table.kmers_host[r0:r0 + m]
            padded = 0 if alphabet == 'prot' else int((ksub >> np.uint64(58) != 0).sum())
            nmax = 10 if A1 == 5 else 3
            big_rows = sub.astype(np.uint64).sum(axis=1) > nmax                      # [G, m]: rows the rank coding escapes
            esc12 = int(((sub != 0) & big_rows[:, None, :]).sum())
            for wire in (8, 4, 12) + (() if alphabet == 'prot' else (8 | 16, 4 | 16, 12 | 16)):
                buf, esc, got = table.compact_chunk(r0, m, wire=wire)
                assert got == wire and buf.numel() == table.compact_bytes(m, wire)
                k, c = _decode_compact(buf, esc, m, lag, alphabet, G, A1, wire)
                assert np.array_equal(k, ksub)
                assert np.array_equal(c, sub)
                want_esc = esc12 if wire & 15 == 12 else int((sub >= (255 if wire & 15 == 8 else 15)).sum())
                assert esc.shape[0] == want_esc + (padded if wire & 16 else 0)
            # the chooser takes the variant with the fewest bytes on the wire, escapes (12 B each) included
            bytes8 = sub.size + 12 * int((sub >= 255).sum())
            bytes4 = sub.size // 2 + 12 * int((sub >= 15).sum())
            want = 4 if bytes4 < bytes8 else 8
            if m * G * 3 // 2 + 12 * esc12 < min(bytes4, bytes8):
                want = 12
            if alphabet != 'prot':
                saved = (2 * lag + 6 + 7) // 8 - (2 * lag + 7) // 8
                if saved > 0 and 12 * padded < m * saved:
                    want |= 16
            assert table.compact_chunk(r0, m)[2] == want


def test_bench_reference_arm_prints_one_json_line():
    """bench.py --impl reference (the CPU arm: the oracle port on the host cores) prints exactly one line on stdout,
    the JSON result with the contract's keys; under torchrun only rank 0 prints."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, os.path.join(root, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '1',
           '--cpu-rows', '2048']
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, check=True).stdout
    lines = [l for l in out.splitlines() if l.strip()]
    assert len(lines) == 1
    line = json.loads(lines[0])
    for key in ('impl', 'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better',
                'scaling', 'dtype', 'data', 'config', 'cpu_baseline', 'e2e'):
        assert key in line, key
    assert line['impl'] == 'reference' and line['cpu_baseline']['kind'] == 'port' and line['value'] > 0
    env = dict(os.environ, RANK='1', WORLD_SIZE='2', LOCAL_RANK='1')
    assert subprocess.run(cmd, capture_output=True, text=True, timeout=300, check=True, env=env).stdout.strip() == ''
