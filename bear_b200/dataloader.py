"""Packed, device-resident k-mer count tables and the BMM marginal likelihood.

Mirrors the reference's ``bear_model/dataloader.py`` API (``dataloader``, ``sparse_dataloader``,
``bmm_likelihood``), but instead of a ``tf.data`` pipeline of byte strings and float64 count tensors
(dataloader.py:36-46, 82-105) the text is parsed ONCE by the C++ packer into
``uint64`` 2-bit packed k-mers + group-planar ``uint32`` counts, uploaded, and kept resident in HBM.
The fused kernels stream that table directly; iterating a dataset still yields the reference's
``(kmers, counts[B, G, A+1])`` batches (unpacked on the device) for API compatibility.
"""
import ctypes
import os

import numpy as np
import torch

from . import _lib
from ._lib import lib, check, ptr

ALPHABET_SIZES = {'dna': 4, 'rna': 4, 'prot': 20}


def _round_up(x, m):
    return (x + m - 1) // m * m


class KmerBatch:
    """The k-mer half of a batch: packed codes on the device; ``.numpy()`` gives the byte strings the
    reference's ``kmers.numpy()`` would (dataloader.py:46)."""

    def __init__(self, packed, lag, alphabet):
        self.packed, self.lag, self.alphabet = packed, lag, alphabet

    def __len__(self):
        return int(self.packed.shape[0])

    def numpy(self):
        return decode_kmers(self.packed.cpu().numpy().view(np.uint64), self.lag, self.alphabet)


def encode_kmers(kmers, alphabet):
    """list / array of equal-length k-mer strings (or bytes) -> uint64 packed codes (numpy), lag."""
    kmers = [k.decode() if isinstance(k, (bytes, np.bytes_)) else str(k) for k in np.asarray(kmers).ravel()]
    n = len(kmers)
    lag = len(kmers[0]) if n else 1
    if any(len(k) != lag for k in kmers):
        raise ValueError('k-mers must all have the same length')
    out = np.empty(n, dtype=np.uint64)
    text = ''.join(kmers).encode()
    check(lib.bear_encode_kmers(text, n, lag, _lib.ALPHABET_IDS[alphabet], ptr(out)))
    return out, lag


def decode_kmers(packed, lag, alphabet):
    packed = np.ascontiguousarray(packed, dtype=np.uint64)
    buf = np.empty(packed.size * lag, dtype=np.uint8)
    check(lib.bear_decode_kmers(ptr(packed), packed.size, lag, _lib.ALPHABET_IDS[alphabet], ptr(buf)))
    return buf.view('S%d' % lag) if packed.size else np.empty(0, dtype='S%d' % lag)


class KmerTable:
    """A packed table: ``kmers`` uint64 [K] and ``counts`` uint32 [G, A1, stride] (host numpy arrays,
    uploaded lazily).  ``stride`` is the plane pitch (multiple of 4 elements)."""

    def __init__(self, kmers, counts, num_rows, lag, alphabet):
        self.kmers_host = kmers
        self.counts_host = counts
        self.num_rows = int(num_rows)
        self.lag = int(lag)
        self.alphabet = alphabet
        self.num_ds = int(counts.shape[0])
        self.A1 = int(counts.shape[1])
        self.stride = int(counts.shape[2])
        self._dev = None
        self._sorted = None
        self.skipped_rows = 0          # rows dropped by from_file(on_invalid='skip')

    # -- construction -------------------------------------------------------------------------
    @classmethod
    def from_file(cls, file, alphabet, num_ds, sparse=False, header=None, on_invalid='error'):
        """``on_invalid``: what a DNA / RNA k-mer with a symbol outside the alphabet (e.g. 'N') does -- 'error' (default)
        or 'skip': the row is dropped and counted in ``table.skipped_rows`` (the reference one-hots such symbols to
        zero rows, core.py:162; 2-bit codes cannot hold them)."""
        if on_invalid not in ('error', 'skip'):
            raise ValueError("on_invalid must be 'error' or 'skip'")
        if on_invalid == 'skip':
            old = check(lib.bear_pack_set_invalid_policy(1))
            try:
                table = cls.from_file(file, alphabet, num_ds, sparse=sparse, header=header)
            finally:
                lib.bear_pack_set_invalid_policy(old)
            bad = table.kmers_host[:table.num_rows] == np.uint64(0xffffffffffffffff)
            nbad = int(bad.sum())
            if nbad:
                table = table.take(np.flatnonzero(~bad))
            table.skipped_rows = nbad
            return table
        if header is None:
            header = bool(sparse)
        path = os.fsencode(file)
        K = lib.bear_count_rows(path, int(header))
        check(K)
        A1 = ALPHABET_SIZES[alphabet] + 1
        stride = max(_round_up(K, 4), 4)
        kmers = np.zeros(stride, dtype=np.uint64)
        counts = np.zeros((num_ds, A1, stride), dtype=np.uint32)
        rows, lag = ctypes.c_int64(0), ctypes.c_int(0)
        fn = lib.bear_pack_sparse if sparse else lib.bear_pack_tsv
        check(fn(path, int(header), _lib.ALPHABET_IDS[alphabet], num_ds, 0, K, ptr(kmers), ptr(counts), stride,
                 ctypes.byref(rows), ctypes.byref(lag)))
        return cls(kmers, counts, rows.value, max(lag.value, 1), alphabet)

    @classmethod
    def from_file_shard(cls, file, alphabet, num_ds, batch_size, rank, world, sparse=False, header=None, row_offset=0,
                        dataset_rows=None):
        """This rank's rows of a count file: of every global batch of ``batch_size`` consecutive rows the contiguous
        slice ``KmerDataset.shard`` would give it, parsed straight from the file -- no rank ever holds (or parses the
        numbers of) the whole table.  ``row_offset`` / ``dataset_rows``: the file is one of several of a dataset and
        holds its rows [row_offset, row_offset + rows of the file) of ``dataset_rows``.  Returns (table, rows of the
        file)."""
        if header is None:
            header = bool(sparse)
        path = os.fsencode(file)
        K = check(lib.bear_count_rows(path, int(header)))
        total = K if dataset_rows is None else int(dataset_rows)
        A1 = ALPHABET_SIZES[alphabet] + 1
        local = rank_rows_before(row_offset + K, total, batch_size, rank, world) - rank_rows_before(row_offset, total, batch_size, rank, world)
        stride = max(_round_up(local, 4), 4)
        kmers = np.zeros(stride, dtype=np.uint64)
        counts = np.zeros((num_ds, A1, stride), dtype=np.uint32)
        rows, lag = ctypes.c_int64(0), ctypes.c_int(0)
        check(lib.bear_pack_shard(path, int(bool(sparse)), int(header), _lib.ALPHABET_IDS[alphabet], num_ds, int(batch_size),
                                  int(world), int(rank), int(row_offset), total, local, ptr(kmers), ptr(counts), stride,
                                  ctypes.byref(rows), ctypes.byref(lag)))
        assert rows.value == local
        if local == 0 and K > 0:                 # a rank without rows still reports the table's lag (first row)
            k1, c1 = np.zeros(4, dtype=np.uint64), np.zeros((num_ds, A1, 4), dtype=np.uint32)
            fn = lib.bear_pack_sparse if sparse else lib.bear_pack_tsv
            check(fn(path, int(header), _lib.ALPHABET_IDS[alphabet], num_ds, 0, 1, ptr(k1), ptr(c1), 4, ctypes.byref(rows),
                     ctypes.byref(lag)))
        return cls(kmers, counts, local, max(lag.value, 1), alphabet), K

    @classmethod
    def from_arrays(cls, kmers, counts, alphabet):
        """kmers: strings or uint64 codes with ``lag`` given as (codes, lag); counts [K, G, A1] integers."""
        if isinstance(kmers, tuple):
            codes, lag = kmers
            codes = np.ascontiguousarray(codes, dtype=np.uint64)
        else:
            codes, lag = encode_kmers(kmers, alphabet)
        counts = np.asarray(counts)
        K, G, A1 = counts.shape
        if A1 != ALPHABET_SIZES[alphabet] + 1:
            raise ValueError('counts last axis must be alphabet_size + 1')
        if np.any(counts < 0) or np.any(counts != np.floor(counts)) or np.any(counts > 4294967295):
            raise ValueError('counts must be integers in [0, 2^32)')
        stride = max(_round_up(K, 4), 4)
        k = np.zeros(stride, dtype=np.uint64)
        k[:K] = codes
        c = np.zeros((G, A1, stride), dtype=np.uint32)
        c[:, :, :K] = np.transpose(counts.astype(np.uint32), (1, 2, 0))
        return cls(k, c, K, lag, alphabet)

    @classmethod
    def from_device(cls, kmers_dev, counts_dev, num_rows, lag, alphabet):
        """Adopt device tensors (int64 [stride], int32 [G, A1, stride]) produced on the GPU
        (synthetic benchmark tables); no host copy exists."""
        self = cls.__new__(cls)
        self.kmers_host = self.counts_host = None
        self.num_rows, self.lag, self.alphabet = int(num_rows), int(lag), alphabet
        self.num_ds, self.A1, self.stride = (int(s) for s in counts_dev.shape)
        self._dev = (kmers_dev, counts_dev)
        self._sorted = None
        return self

    @classmethod
    def concat(cls, tables):
        t0 = tables[0]
        K = sum(t.num_rows for t in tables)
        stride = max(_round_up(K, 4), 4)
        kmers = np.zeros(stride, dtype=np.uint64)
        counts = np.zeros((t0.num_ds, t0.A1, stride), dtype=np.uint32)
        o = 0
        for t in tables:
            if (t.lag, t.alphabet, t.num_ds) != (t0.lag, t0.alphabet, t0.num_ds):
                raise ValueError('tables differ in lag / alphabet / num_ds')
            kmers[o:o + t.num_rows] = t.kmers_host[:t.num_rows]
            counts[:, :, o:o + t.num_rows] = t.counts_host[:, :, :t.num_rows]
            o += t.num_rows
        return cls(kmers, counts, K, t0.lag, t0.alphabet)

    # -- packed binary shard cache ----------------------------------------------------------------
    MAGIC = b'BEARPACK1\n'

    def save(self, path):
        """Write the packed table as one binary file (header + raw uint64 k-mers + raw uint32 counts) so that
        later runs skip the text parse; ``KmerTable.load`` maps it back."""
        import json
        if self.kmers_host is None:
            k, c = self._dev
            kmers, counts = k.cpu().numpy().view(np.uint64), c.cpu().numpy().view(np.uint32)
        else:
            kmers, counts = self.kmers_host, self.counts_host
        meta = json.dumps({'num_rows': self.num_rows, 'lag': self.lag, 'alphabet': self.alphabet, 'num_ds': self.num_ds,
                           'A1': self.A1, 'stride': self.stride}).encode()
        with open(path, 'wb') as fh:
            fh.write(self.MAGIC)
            fh.write(len(meta).to_bytes(8, 'little'))
            fh.write(meta)
            fh.write(b'\0' * (-fh.tell() % 64))
            np.ascontiguousarray(kmers).tofile(fh)
            np.ascontiguousarray(counts).tofile(fh)

    @classmethod
    def load(cls, path):
        import json
        with open(path, 'rb') as fh:
            if fh.read(len(cls.MAGIC)) != cls.MAGIC:
                raise ValueError('%s is not a BEARPACK file' % path)
            meta = json.loads(fh.read(int.from_bytes(fh.read(8), 'little')))
            off = fh.tell() + (-fh.tell() % 64)
        stride, G, A1 = meta['stride'], meta['num_ds'], meta['A1']
        kmers = np.memmap(path, dtype=np.uint64, mode='r', offset=off, shape=(stride,))
        counts = np.memmap(path, dtype=np.uint32, mode='r', offset=off + 8 * stride, shape=(G, A1, stride))
        return cls(kmers, counts, meta['num_rows'], meta['lag'], meta['alphabet'])

    def take(self, index):
        """Rows ``index`` (numpy int array) as a new host table."""
        index = np.asarray(index, dtype=np.int64)
        K = index.size
        stride = max(_round_up(K, 4), 4)
        kmers = np.zeros(stride, dtype=np.uint64)
        counts = np.zeros((self.num_ds, self.A1, stride), dtype=np.uint32)
        kmers[:K] = self.kmers_host[index]
        counts[:, :, :K] = self.counts_host[:, :, index]
        return KmerTable(kmers, counts, K, self.lag, self.alphabet)

    # -- device residency ---------------------------------------------------------------------
    def compact_chunk(self, r0, n, out=None, wire=None):
        """Rows [r0, r0+n) in the compact transfer format (k-mer byte planes, 8- or 4-bit count planes + escapes for
        what does not fit, see include/bear_b200.h): (uint8 tensor, uint32 escape tensor [n_esc, 3], wire).
        ``wire``: the variant (count bits, + _lib.WIRE_START_ESC); None = the one with the fewest bytes on the wire
        for these rows (bear_compact_choose_wire).  ``out`` = a reusable (pinned) uint8 tensor of at least
        ``compact_bytes(n)`` elements."""
        aid = _lib.ALPHABET_IDS[self.alphabet]
        if wire is None:
            wire = lib.bear_compact_choose_wire(ptr(self.kmers_host), ptr(self.counts_host), self.stride, r0, n,
                                                self.lag, aid, self.num_ds)
            check(wire)
        nbytes = lib.bear_compact_bytes(n, self.lag, aid, self.num_ds, wire)
        check(nbytes)
        buf = out[:nbytes] if out is not None else torch.empty(nbytes, dtype=torch.uint8)
        cap = 1024
        while True:
            esc = np.empty((cap, 3), dtype=np.uint32)
            need = ctypes.c_int64(0)
            check(lib.bear_compact_table(ptr(self.kmers_host), ptr(self.counts_host), self.stride, r0, n, self.lag, aid,
                                         self.num_ds, wire, ctypes.c_void_p(buf.data_ptr()), ptr(esc), cap,
                                         ctypes.byref(need)))
            if need.value <= cap:
                return buf, torch.from_numpy(esc[:need.value].view(np.int32).copy()), wire
            cap = int(need.value)

    def compact_bytes(self, n, wire=8):
        return int(lib.bear_compact_bytes(n, self.lag, _lib.ALPHABET_IDS[self.alphabet], self.num_ds, wire))

    def device_tensors(self):
        """(kmers int64 [stride], counts int32 [G, A1, stride]) on the current CUDA device; the bit
        patterns are the uint64 / uint32 of the packed layout.  The upload crosses the bus in the compact transfer
        format (7.6 - 11 instead of 28 bytes per row for a one-column DNA table at lag 20) through a bounded pinned
        buffer and is expanded on the device (bear_expand_table), bit-exactly."""
        if self._dev is None:
            dev = _lib.device()
            k = torch.empty(self.stride, dtype=torch.int64, device=dev)
            c = torch.zeros((self.num_ds, self.A1, self.stride), dtype=torch.int32, device=dev)
            k[self.num_rows:].zero_()
            aid = _lib.ALPHABET_IDS[self.alphabet]
            step = 1 << 24                                  # multiple of 4: vector stores in the expansion
            stage = torch.empty(self.compact_bytes(min(step, max(self.num_rows, 1))), dtype=torch.uint8).pin_memory()
            dbuf = torch.empty_like(stage, device=dev)
            for lo in range(0, self.num_rows, step):
                n = min(step, self.num_rows - lo)
                buf, esc, bits = self.compact_chunk(lo, n, out=stage)
                dbuf[:buf.numel()].copy_(buf, non_blocking=True)
                desc = esc.to(dev) if esc.numel() else None
                check(lib.bear_expand_table(ptr(dbuf), ptr(desc), esc.shape[0], n, self.lag, aid, self.num_ds, bits,
                                            ptr(k), ptr(c), self.stride, lo, _lib.stream()))
                torch.cuda.current_stream().synchronize()  # the pinned stage is reused by the next chunk
            self._dev = (k, c)
            self._sorted = None
        return self._dev

    def sorted_index(self):
        """(sorted packed codes, row order) of the resident table, built once per upload: every point query
        (get_var_probs.lookup_counts, assemble) is a binary search against it instead of a fresh O(K log K) sort."""
        k, _ = self.device_tensors()
        if getattr(self, '_sorted', None) is None or self._sorted[0].device != k.device:
            keys, order = torch.sort(k[:self.num_rows])
            self._sorted = (keys, order)
        return self._sorted

    def col_ptr(self, ds_loc):
        """Device address of plane 0 of count column ``ds_loc``."""
        _, c = self.device_tensors()
        if not 0 <= ds_loc < self.num_ds:
            raise IndexError('count column %d out of range (num_ds=%d)' % (ds_loc, self.num_ds))
        return ctypes.c_void_p(c.data_ptr() + 4 * ds_loc * self.A1 * self.stride)

    def kmers_str(self, row0=0, n=None):
        n = self.num_rows - row0 if n is None else n
        if self.kmers_host is None:
            host = self._dev[0][row0:row0 + n].cpu().numpy().view(np.uint64)
        else:
            host = self.kmers_host[row0:row0 + n]
        return decode_kmers(host, self.lag, self.alphabet)


def rank_rows_before(g, K, batch_size, rank, world):
    """Rows owned by ``rank`` among rows [0, g) of a K-row dataset cut into global batches of ``batch_size`` rows
    (the arithmetic of bear_pack_shard)."""
    b, i = divmod(int(g), int(batch_size))
    full_per = -(-int(batch_size) // world)
    cnt_full = max(min((rank + 1) * full_per, batch_size) - min(rank * full_per, batch_size), 0)
    n_b = min(batch_size, K - b * batch_size)
    in_b = 0
    if n_b > 0:
        per = -(-n_b // world)
        cnt_b = max(min((rank + 1) * per, n_b) - min(rank * per, n_b), 0)
        in_b = min(max(i - rank * per, 0), cnt_b)
    return b * cnt_full + in_b


def shard_ranges(K, batch_size, rank, world):
    """How a table of K rows cut into global batches of ``batch_size`` rows spreads over ``world`` ranks: rank keeps
    the contiguous slice [rank * per, (rank + 1) * per) of every batch, per = ceil(rows of the batch / world).
    Returns (local (row0, n) per batch, global rows per batch, global index of the first local row per batch)."""
    ranges, grows, ids, o = [], [], [], 0
    for r0 in range(0, K, batch_size):
        n = min(batch_size, K - r0)
        per = -(-n // world)
        lo, hi = min(rank * per, n), min((rank + 1) * per, n)
        ranges.append((o, hi - lo))
        grows.append(n)
        ids.append(r0 + lo)
        o += hi - lo
    return ranges, grows, ids


class KmerDataset:
    """What ``dataloader`` returns: a packed table cut into minibatches of ``batch_size`` consecutive
    rows (dataloader.py:37), optionally repeated (``.repeat(epochs)``, models/train_bear_net.py:87).

    Under ``torch.distributed`` each rank keeps only its contiguous slice of every global batch
    (``shard(rank, world)``); the fused kernels see the local slices, ``global_batch_rows`` keeps the
    reference's ``num_kmers / batch`` factor (bear_net.py:190) independent of the GPU count.
    """

    def __init__(self, table, batch_size, repeats=1, ranges=None, global_rows=None, map_fn=None, row_ids=None):
        self.table = table
        self.batch_size = int(batch_size)
        self.repeats = int(repeats)
        K = table.num_rows
        if ranges is None:
            ranges = [(r0, min(self.batch_size, K - r0)) for r0 in range(0, K, self.batch_size)]
            global_rows = [n for _, n in ranges]
        self.ranges = ranges                  # local (row0, n) per batch
        self.global_rows = global_rows        # rows of the same batch summed over all ranks
        # index, in the whole (unsharded) table, of the first local row of every batch: keys the argmax tie-break
        # noise of the evaluation, so that results do not depend on the number of ranks
        self.row_ids = [r0 for r0, _ in ranges] if row_ids is None else row_ids
        self.map_fn = map_fn

    # tf.data-like surface used by the reference scripts
    def repeat(self, count):
        return KmerDataset(self.table, self.batch_size, self.repeats * int(count), self.ranges, self.global_rows, self.map_fn,
                           self.row_ids)

    def cache(self):
        return self

    def prefetch(self, n):
        return self

    def map(self, fn, num_parallel_calls=None):
        prev = self.map_fn
        f = fn if prev is None else (lambda *a: fn(*_as_tuple(prev(*a))))
        return KmerDataset(self.table, self.batch_size, self.repeats, self.ranges, self.global_rows, f, self.row_ids)

    def __len__(self):
        return len(self.ranges) * self.repeats

    def batches(self):
        """(row0, n_local, n_global) for every step, repeats included."""
        for _ in range(self.repeats):
            for (r0, n), g in zip(self.ranges, self.global_rows):
                yield r0, n, g

    def shard(self, rank, world):
        """This rank's contiguous slice of every batch, gathered into a table of its own."""
        if world == 1:
            return self
        if self.table.kmers_host is None:
            raise ValueError('shard() needs a host table; device-generated tables are built per rank')
        idx, ranges, grows, ids, o = [], [], [], [], 0
        for (r0, n), gid in zip(self.ranges, self.row_ids):
            per = -(-n // world)
            lo, hi = min(rank * per, n), min((rank + 1) * per, n)
            idx.append(np.arange(r0 + lo, r0 + hi))
            ranges.append((o, hi - lo))
            grows.append(n)
            ids.append(gid + lo)
            o += hi - lo
        local = self.table.take(np.concatenate(idx) if idx else np.zeros(0, np.int64))
        return KmerDataset(local, self.batch_size, self.repeats, ranges, grows, self.map_fn, ids)

    def __iter__(self):
        """Yields the reference's batch elements: (KmerBatch, counts float64 [B, G, A1] on the device),
        passed through ``map`` functions if any."""
        t = self.table
        k, c = t.device_tensors()
        for r0, n, _ in self.batches():
            out = torch.empty((n, t.num_ds, t.A1), dtype=torch.float64, device=k.device)
            check(lib.bear_unpack_counts(ptr(c), t.stride, r0, n, t.num_ds, t.A1, ptr(out), _lib.stream()))
            item = (KmerBatch(k[r0:r0 + n], t.lag, t.alphabet), out)
            yield item if self.map_fn is None else self.map_fn(*item)


def _as_tuple(x):
    return x if isinstance(x, tuple) else (x,)


def _local_shard(ds):
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return ds.shard(dist.get_rank(), dist.get_world_size())
    return ds


def _load_file(file, alphabet, batch_size, num_ds, sparse, header):
    """One count file -> this process's KmerDataset.  Under ``torch.distributed`` every rank parses only its own
    slice of every global batch (``bear_pack_shard``); nothing holds the whole table."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        rank, world = dist.get_rank(), dist.get_world_size()
        table, K = KmerTable.from_file_shard(file, alphabet, num_ds, batch_size, rank, world, sparse=sparse, header=header)
        ranges, grows, ids = shard_ranges(K, int(batch_size), rank, world)
        return KmerDataset(table, batch_size, 1, ranges, grows, None, ids)
    return KmerDataset(KmerTable.from_file(file, alphabet, num_ds, sparse=sparse, header=header), batch_size)


def dataloader(file, alphabet, batch_size, num_ds, cache=True, header=False, n_par=1, dtype=torch.float64):
    """Dense TSV ``kmer \\t [[counts g0],[counts g1],...]`` -> KmerDataset (reference:
    dataloader.py:6-50).  ``cache`` / ``n_par`` are accepted for signature parity; the packed table is
    always resident."""
    return _load_file(file, alphabet, batch_size, num_ds, False, header)


def sparse_dataloader(file, alphabet, batch_size, num_ds, cache=False, header=True, n_par=1, dtype=torch.float64):
    """Sparse ``kmer; [[g,b],...]; [v,...]`` file -> KmerDataset (reference: dataloader.py:52-109)."""
    return _load_file(file, alphabet, batch_size, num_ds, True, header)


def load_files(files, alphabet, batch_size, num_ds, sparse=False, header=None):
    """Several files of one dataset (models/train_bear_net.py:78-86 interleaves them; here they are
    concatenated in name order into one resident table)."""
    import torch.distributed as dist
    files = sorted(files)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        # every rank parses only its slice of every global batch of the concatenated dataset
        rank, world = dist.get_rank(), dist.get_world_size()
        hdr = bool(sparse) if header is None else header
        rows = [count_rows(f, header=hdr) for f in files]
        K, parts, off = sum(rows), [], 0
        for f, n in zip(files, rows):
            parts.append(KmerTable.from_file_shard(f, alphabet, num_ds, batch_size, rank, world, sparse=sparse, header=header,
                                                   row_offset=off, dataset_rows=K)[0])
            off += n
        table = parts[0] if len(parts) == 1 else KmerTable.concat(parts)
        ranges, grows, ids = shard_ranges(K, int(batch_size), rank, world)
        return KmerDataset(table, batch_size, 1, ranges, grows, None, ids)
    tables = [KmerTable.from_file(f, alphabet, num_ds, sparse=sparse, header=header) for f in files]
    table = tables[0] if len(tables) == 1 else KmerTable.concat(tables)
    return KmerDataset(table, batch_size)


def pack_files(files, out_path, alphabet, num_ds, sparse=False, header=None):
    """Parse text count files once (multi-threaded C++ packer) and write a BEARPACK binary shard."""
    tables = [KmerTable.from_file(f, alphabet, num_ds, sparse=sparse, header=header) for f in sorted(files)]
    table = tables[0] if len(tables) == 1 else KmerTable.concat(tables)
    table.save(out_path)
    return table


def load_packed(path, batch_size):
    """KmerDataset over a BEARPACK shard written by ``pack_files`` / ``KmerTable.save``."""
    return _local_shard(KmerDataset(KmerTable.load(path), batch_size))


def count_rows(file, header=False):
    """`wc -l` of models/train_bear_net.py:54-55 (non-empty data rows)."""
    return check(lib.bear_count_rows(os.fsencode(file), int(header)))


def _allreduce_sum(t):
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def bmm_likelihood(data, alpha, dtype=torch.float64):
    """BMM marginal log-likelihood of every count column for every prior in ``alpha`` -> [num_ds, V]
    (reference: dataloader.py:111-147).  ``data`` is a KmerDataset (a ``.map(lambda k, c: c)`` view of
    it is accepted, as in the reference's usage); the packed table is streamed once by
    ``bear_bmm_likelihood``, and the [G, V] partial sums are allreduced once at the end."""
    if not isinstance(data, KmerDataset):
        raise TypeError('bmm_likelihood needs a KmerDataset (from dataloader / sparse_dataloader)')
    t = data.table
    k, c = t.device_tensors()
    alpha = torch.as_tensor(np.asarray(alpha, dtype=np.float64)).reshape(-1)
    V = alpha.numel()
    out = torch.zeros((t.num_ds, V), dtype=torch.float64, device=k.device)
    ws = torch.empty(lib.bear_workspace_doubles(t.num_rows, t.lag, 0), dtype=torch.float64, device=k.device)
    for v0 in range(0, V, _lib.MAX_MODELS):
        a = alpha[v0:v0 + _lib.MAX_MODELS].to(k.device)
        part = torch.zeros((t.num_ds, a.numel()), dtype=torch.float64, device=k.device)
        for r0, n, _ in data.batches():
            check(lib.bear_bmm_likelihood(ptr(c), t.stride, r0, n, t.num_ds, t.A1, ptr(a), a.numel(), ptr(part),
                                          ptr(ws), _lib.stream()))
        out[:, v0:v0 + a.numel()] = part
    return _allreduce_sum(out).cpu()
