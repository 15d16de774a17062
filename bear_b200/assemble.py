"""Sequence generation from a BEAR / BMM model, extending seed sequences in both directions.

Mirrors the reference's ``bear_model/assemble.py`` (``assemble_no_ends`` :21-184): for every generated
sequence ONE autoregressive model is drawn from the BEAR posterior (a Dirichlet draw per k-mer, the same
draw whenever the sequence meets that k-mer again, assemble.py:109-131) and the sequence is extended
letter by letter with Gumbel-max sampling over the non-stop letters (:133); the left flank is generated
on the reverse complement (:81,139) and flipped back.  The transition probabilities come from
``get_var_probs.get_pdf`` (device: concentrations, log-Gamma draws, normalisation).

The reference reads the transition counts from a KMC database through ``py_kmc_api``
(``get_var_probs.make_kmc_genome_counter``, assemble.py:69), which is not available here: the counts come
from a resident k-mer table (``data``, a ``dataloader.KmerDataset``) through one packed-code join per
step (``get_var_probs.lookup_counts``).  ``reverse=True`` adds the counts of the reverse-complement
(k+1)-mers, like the KMC counter's ``reverse`` flag; pass ``reverse=False`` for a table that was already
summarised with ``-r``.  Passing ``kmc_path`` raises.
"""
import os

import numpy as np

from . import core, get_var_probs

_COMP = {'dna': str.maketrans('ACGT', 'TGCA'), 'rna': str.maketrans('ACGU', 'UGCA')}


def reverse_complement(seq, alphabet_name='dna'):
    """Reverse complement of a nucleotide string (the reference uses Bio.Seq.reverse_complement, assemble.py:81)."""
    return seq.translate(_COMP[alphabet_name])[::-1]


def read_fasta(path):
    """Sequences of a FASTA file, in file order (assemble.py:75 uses Bio.SeqIO.parse)."""
    seqs, cur = [], None
    with open(path) as fh:
        for line in fh:
            line = line.strip()
            if line.startswith('>'):
                if cur is not None:
                    seqs.append(''.join(cur))
                cur = []
            elif line and cur is not None:
                cur.append(line)
    if cur is not None:
        seqs.append(''.join(cur))
    return seqs


def sitewise_entropy(seqs, alphabet_name):
    """Entropy of the letter distribution at every site of equally long sequences (assemble.py:144-146)."""
    letters = core.alphabets_en[alphabet_name][:-1]
    index = {ch: i for i, ch in enumerate(letters)}
    n, length = len(seqs), len(seqs[0])
    freq = np.zeros((length, len(letters) + 1))
    for s in seqs:
        assert len(s) == length
        for i, ch in enumerate(s):
            if ch in index:                      # symbols outside the alphabet one-hot to a zero row (core.py:162)
                freq[i, index[ch]] += 1.0
    p = freq / n
    with np.errstate(divide='ignore', invalid='ignore'):
        return -np.where(p > 0, p * np.log(p), 0.0).sum(-1)


class TableCounter:
    """counter(kmers) -> [K, A+1] transition counts out of the query k-mers, from one column of a resident
    table; the stand-in for the reference's KMC genome counter (get_var_probs.py:196-289, ``no_end=True``:
    the stop column is not used by the generator)."""

    def __init__(self, data, alphabet_name, train_col=0, reverse=True):
        if reverse and alphabet_name not in ('dna', 'rna'):
            raise ValueError('reverse-complement counts need a nucleotide alphabet')
        self.data, self.alphabet_name, self.col, self.reverse = data, alphabet_name, train_col, reverse
        self.letters = core.alphabets_en[alphabet_name][:-1]

    def __call__(self, kmers):
        kmers = [str(k) for k in kmers]
        A = len(self.letters)
        if not kmers:
            return np.zeros((0, A + 1))
        counts, _ = get_var_probs.lookup_counts(self.data, np.array(kmers), self.alphabet_name)
        out = counts[:, self.col, :].cpu().numpy().copy()
        if self.reverse:
            # the reverse complement of kmer+b is comp(b)+rc(kmer): a transition out of the k-mer
            # comp(b)+rc(kmer)[:-1] to the letter comp(kmer[0])
            comp = {ch: ch.translate(_COMP[self.alphabet_name]) for ch in self.letters}
            index = {ch: i for i, ch in enumerate(self.letters)}
            queries, where = [], []
            for i, k in enumerate(kmers):
                if any(ch not in index for ch in k):
                    continue                      # start-padded k-mers have no reverse strand
                rc = reverse_complement(k, self.alphabet_name)
                for b, ch in enumerate(self.letters):
                    queries.append(comp[ch] + rc[:-1])
                    where.append((i, b, index[comp[k[0]]]))
            if queries:
                rc_counts, _ = get_var_probs.lookup_counts(self.data, np.array(queries), self.alphabet_name)
                rc_counts = rc_counts[:, self.col, :].cpu().numpy()
                w = np.array(where)
                np.add.at(out, (w[:, 0], w[:, 1]), rc_counts[np.arange(len(queries)), w[:, 2]])
        return out


def _extend(seqs, length_to_gen, lag, letters, counter, h, ar_func, vans, alphabet_name, get_map, rng, seed):
    """Extends every sequence of one batch to the right by length_to_gen[i] letters (assemble.py:86-140)."""
    S = len(seqs)
    target = np.asarray(length_to_gen, dtype=np.int64)
    new_seq = [[] for _ in range(S)]
    done = np.zeros(S, dtype=np.int64)
    end = [s[-lag:] for s in seqs]
    for s, t in zip(end, target):
        if t > 0 and len(s) < lag:
            raise ValueError('seed sequences must be at least lag = %d letters long' % lag)
    index, all_pdf, calls = {}, None, 0
    active = np.flatnonzero(done < target)
    while active.size:
        cur = [end[i] for i in active]
        fresh = sorted({k for k in cur if k not in index})
        if fresh:
            counts = counter(fresh)[:, None, :]
            # one posterior draw per (k-mer, sequence of the batch): column s is sequence s's model
            pdf = get_var_probs.get_pdf(np.array(fresh), counts, h, ar_func, S, vans, 0, alphabet_name, get_map,
                                        output='numpy', seed=None if seed is None else seed + calls)[:, :-1, 0, :]
            calls += 1
            for i, k in enumerate(fresh):
                index[k] = (0 if all_pdf is None else all_pdf.shape[0]) + i
            all_pdf = pdf if all_pdf is None else np.concatenate([all_pdf, pdf])
        rows = np.array([index[k] for k in cur])
        logp = all_pdf[rows, :, 0] if get_map else all_pdf[rows, :, active]              # [active, A]
        pick = np.argmax(rng.gumbel(size=logp.shape) + logp, axis=-1)                     # assemble.py:133
        for i, b in zip(active, pick):
            new_seq[i].append(letters[b])
            end[i] = end[i][1:] + letters[b]
            done[i] += 1
        active = np.flatnonzero(done < target)
    return [''.join(s) for s in new_seq]


def assemble_no_ends(seqs_fa_file, lengths_to_gen, num_to_gen, bear_path, kmc_path=None,
                     h=None, reverse=True, save_folder=None, batch_size=100,
                     van=None, lag=None, alphabet_name=None, get_map=False,
                     data=None, train_col=0, seed=None):
    """Generate sequences from seeds with a BEAR model (or a BMM when ``van`` is given) that cannot stop
    (assemble.py:21-184; same arguments, plus ``data`` / ``train_col`` for the count table and ``seed``).

    seqs_fa_file : FASTA file of seed sequences (or a list of str).
    lengths_to_gen : [len(seqs), 2] letters to generate backwards / forwards of every seed.
    num_to_gen : sequences generated per seed.
    Returns (gen_seqs [len(seqs), num_to_gen] array of str, sw_ent list of per-site entropies)."""
    if kmc_path is not None:
        raise NotImplementedError('the KMC random-access counter needs py_kmc_api; pass the count table as data=')
    ar_func = None
    if bear_path is not None:
        lag, alphabet_name, h_bear, ar_func, table = get_var_probs.load_bear(bear_path)
        if data is None:
            data = table
        if h is None:
            h = h_bear
    if van is not None:
        assert lag is not None and alphabet_name is not None
        vans, ar_func, h = van * np.ones(1), None, None
    else:
        assert ar_func is not None, 'either bear_path or van (with lag, alphabet_name, data) is needed'
        vans, h = [], np.array([h])
    assert data is not None
    letters = core.alphabets_en[alphabet_name][:-1]
    counter = TableCounter(data, alphabet_name, train_col, reverse)
    rng = np.random.default_rng(seed)

    seeds = list(seqs_fa_file) if not isinstance(seqs_fa_file, (str, os.PathLike)) else read_fasta(seqs_fa_file)
    lengths = np.asarray(lengths_to_gen, dtype=np.int64).reshape(len(seeds), 2)
    fwd_seqs = [s for s in seeds for _ in range(num_to_gen)]
    lens_rep = np.repeat(lengths, num_to_gen, axis=0)                                    # [S, 2]
    nucleotide = alphabet_name in ('dna', 'rna')
    if not nucleotide and lens_rep[:, 0].any():
        raise ValueError('backward generation runs on the reverse complement: nucleotide alphabets only')
    rev_seqs = [reverse_complement(s, alphabet_name) for s in fwd_seqs] if nucleotide else fwd_seqs

    flanks = []
    for side, seqs_side in enumerate((rev_seqs, fwd_seqs)):
        out = []
        for lo in range(0, len(seqs_side), batch_size):
            sub_seed = None if seed is None else seed * 7919 + side * 104729 + lo
            out += _extend(seqs_side[lo:lo + batch_size], lens_rep[lo:lo + batch_size, side], lag, letters, counter,
                           h, ar_func, vans, alphabet_name, get_map, rng, sub_seed)
        flanks.append(out)
    gen = [(reverse_complement(left, alphabet_name) if nucleotide else left) + mid + right
           for left, mid, right in zip(flanks[0], fwd_seqs, flanks[1])]
    gen_seqs = np.array(gen, dtype=object).reshape(len(seeds), num_to_gen) if gen else np.zeros((0, num_to_gen), object)
    sw_ent = [sitewise_entropy(list(row), alphabet_name) for row in gen_seqs]

    if save_folder is not None:
        os.makedirs(save_folder, exist_ok=True)
        with open(os.path.join(save_folder, 'seqs.fa'), 'w') as fh:
            for i, row in enumerate(gen_seqs):
                for j, s in enumerate(row):
                    fh.write('>seq{}_rep{}\n{}\n'.format(i, j, s))
        _plot_entropy(sw_ent, lengths, len(letters), save_folder)
    return gen_seqs, sw_ent


def _plot_entropy(sw_ent, lengths, alphabet_size, save_folder):
    """entropy.png / entropy_zoom.png (assemble.py:155-183); entropy.txt when matplotlib is missing."""
    try:
        import matplotlib
        matplotlib.use('Agg')
        from matplotlib import pyplot as plt
    except Exception:
        with open(os.path.join(save_folder, 'entropy.txt'), 'w') as fh:
            for ent in sw_ent:
                fh.write(' '.join('%.6g' % e for e in ent) + '\n')
        return
    plt.figure(figsize=[10, 5])
    plt.xlabel('position')
    plt.ylabel('entropy')
    lo, hi = 0, 0
    for ent, (back, _) in zip(sw_ent, lengths):
        xs = np.arange(len(ent)) - back
        lo, hi = min(lo, xs.min()), max(hi, xs.max())
        plt.plot(xs, ent, color='blue', linewidth=1, alpha=0.1)
    plt.plot([lo, hi], np.log(alphabet_size) * np.ones(2), color='black', linewidth=2)
    plt.xlim([lo, hi])
    plt.ylim([0, plt.ylim()[1]])
    plt.savefig(os.path.join(save_folder, 'entropy.png'), dpi=200)
    plt.xlim([-10, 0])
    plt.savefig(os.path.join(save_folder, 'entropy_zoom.png'), dpi=200)
    plt.close()
