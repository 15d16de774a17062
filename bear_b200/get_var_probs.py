"""Posterior-predictive transition probabilities under BEAR / BMM models, and scores of variants and
whole sequences built from them.

Mirrors the reference's ``bear_model/get_var_probs.py`` public functions (``load_ds`` :35-57,
``load_bear`` :59-82, ``get_pdf`` :91-194, ``parse_var`` :336-341, ``get_bear_probs`` :343-454,
``get_bear_probs_seqs`` :510-631).  The numeric core of ``get_pdf`` runs on the device: the
concentrations are assembled there, MC samples come from ``bear_loggamma_sample`` +
``bear_log_normalize`` (reference: log_gamma.log_gamma minus logsumexp, get_var_probs.py:174-175),
the closed-form marginal from ``bear_dm_logprob`` (get_var_probs.py:162-169).  Instead of scanning
the whole dataset batch by batch with ``np.isin`` on strings (get_var_probs.py:428-451), the query
k-mers are packed and joined against the resident table's packed codes in one device pass; k-mers
absent from the table get zero counts, exactly what the reference's "unseen k-mers" branch does.
The KMC random-access counter (get_var_probs.py:196-289) needs ``py_kmc_api`` and is not part of this
hot path: passing ``kmc_path`` raises.
"""
import configparser
import json
import os

import numpy as np
import torch

from . import _lib, ar_funcs, bear_net, core, dataloader, log_gamma
from ._lib import lib, check, ptr

epsilon = 1e-7


def cross_str_arrays(array1, array2, exch='X'):
    """All concatenations a + b, a-major order (get_var_probs.py:23-33)."""
    return np.array([str(a) + str(b) for a in array1 for b in array2])


def load_ds(files_path, start_token, kmer_batch_size, sparse, alphabet, num_ds, dtype=torch.float64):
    """All files of a dataset as one resident KmerDataset (get_var_probs.py:35-57)."""
    files = [os.path.join(files_path, f) for f in os.listdir(files_path) if f.startswith(start_token)]
    return dataloader.load_files(files, alphabet, kmer_batch_size, num_ds, sparse=sparse)


def load_bear(path, reference_compatible=True):
    """Load a trained model folder (config.cfg + results.pickle) -> (lag, alphabet, h, ar_func, data)
    (get_var_probs.py:59-82).  With ``reference_compatible`` the AR function is wrapped as
    softmax(ar_func(k)) + epsilon like the reference does (a second softmax on already normalised
    probabilities, get_var_probs.py:79-81); pass False to use ar_func(k) + epsilon."""
    import dill
    config = configparser.ConfigParser()
    config.read(os.path.join(path, 'config.cfg'))
    lag = int(config['hyperp']['lag'])
    alphabet = config['data']['alphabet']
    alphabet_size = len(core.alphabets_tf[alphabet]) - 1
    make_ar_func = getattr(ar_funcs, 'make_ar_func_' + config['model']['ar_func_name'])
    af_kwargs = json.loads(config['model']['af_kwargs'])
    with open(os.path.join(path, 'results.pickle'), 'rb') as fh:
        params_restart = dill.load(fh)['params']
    params, h_signed, ar_func = bear_net.change_scope_params(lag, alphabet_size, make_ar_func, af_kwargs, params_restart)
    h = float(np.exp(h_signed.cpu().numpy()))
    data = load_ds(config['data']['files_path'], config['data']['start_token'], int(float(config['train']['batch_size'])),
                   config['data']['sparse'] == 'True', alphabet, int(config['data']['num_ds']))

    def ar_func_tf(kmers):
        f = ar_func(kmers)
        return (torch.softmax(f, dim=-1) if reference_compatible else f) + epsilon
    return lag, alphabet, h, ar_func_tf, data


class _PDF:
    """log-probabilities [K, A1, num_models, mc] indexed by (k+1)-mer strings."""

    def __init__(self, kmers, log_probs, alphabet_name, summed):
        self.index = {str(k): i for i, k in enumerate(kmers)}
        self.letters = {ch: b for b, ch in enumerate(core.alphabets_en[alphabet_name])}
        self.lp, self.summed = log_probs, summed

    def rows(self, kp1mers):
        ki = np.array([self.index[k[:-1]] for k in kp1mers], dtype=np.int64)
        bi = np.array([self.letters[k[-1]] for k in kp1mers], dtype=np.int64)
        return self.lp[ki, bi] if len(ki) else np.zeros((0,) + self.lp.shape[2:])

    def __call__(self, kp1mers):
        r = self.rows(list(kp1mers))
        return r.sum(0) if self.summed else r


def df_to_func(df, num_models, mc_samples, summed=True):
    """Function view of a (k+1)-mer indexed DataFrame (get_var_probs.py:84-89)."""
    if summed:
        return lambda kp1mers_ex: np.sum(df.loc[kp1mers_ex].to_numpy().reshape([-1, num_models, mc_samples]), axis=0)
    return lambda kp1mers_ex: df.loc[kp1mers_ex].to_numpy().reshape([-1, num_models, mc_samples])


def _concentrations(kmers, counts_train, h, ar_func, vans, alphabet_name, get_map):
    """concs [num_models, K, A1] on the device (get_var_probs.py:132-153): BEAR models
    ar_func(k)/h_i first, then BMM priors van_j, all plus the training counts; with get_map the raw
    AR probabilities are prepended."""
    dev = _lib.device()
    K, A1 = counts_train.shape
    blocks = []
    ar_vals = None
    if ar_func is not None:
        ar_vals = ar_func(core.tf_one_hot(kmers, alphabet_name)).to(torch.float64)
        hh = torch.as_tensor(np.asarray(h, dtype=np.float64), device=dev).reshape(-1, 1, 1)
        blocks.append(ar_vals[None, :, :] / hh)
    if len(vans) > 0:
        vv = torch.as_tensor(np.asarray(vans, dtype=np.float64), device=dev).reshape(-1, 1, 1)
        blocks.append(vv * torch.ones((1, K, A1), dtype=torch.float64, device=dev))
    concs = torch.cat(blocks, 0) + counts_train[None, :, :]
    if ar_vals is not None and get_map:
        concs = torch.cat([ar_vals[None], concs], 0)
    return concs.contiguous()


def get_pdf(kmers, counts, h, ar_func, mc_samples, vans, train_col, alphabet_name, get_map,
            get_marg=False, summed=True, output='func', seed=None):
    """Probabilities of all (k+1)-mer transitions out of ``kmers`` (get_var_probs.py:91-194).

    kmers : array of str; counts : [K, num_ds, A1] array / tensor; h : array of BEAR h values;
    ar_func : callable on one-hot tensors or None (BMM only); vans : BMM priors.
    get_map -> log(concs / sum concs); get_marg -> a function (kmers, counts) -> summed closed-form
    marginal log-probabilities [num_models]; otherwise ``mc_samples`` posterior draws
    log Dirichlet(concs).  output: 'func' (callable on lists of (k+1)-mers), 'df' (pandas) or 'numpy'
    ([kmer, letter, model, mc])."""
    assert not (get_marg and get_map), "pick marg or map"
    assert not (get_marg and output != 'func'), "not implemented"
    dev = _lib.device()
    kmers = np.asarray(kmers).astype(str)
    if get_map or get_marg:
        mc_samples = 1
    counts = torch.as_tensor(np.asarray(counts.cpu() if isinstance(counts, torch.Tensor) else counts, dtype=np.float64))
    counts_train = counts[:, train_col, :].to(dev).contiguous()
    concs = _concentrations(kmers, counts_train, h, ar_func, vans, alphabet_name, get_map)
    M, K, A1 = concs.shape

    if get_marg:
        index = {k: i for i, k in enumerate(kmers)}

        def prob_func(q_kmers, q_counts):
            sel = torch.as_tensor([index[str(k)] for k in q_kmers], device=dev, dtype=torch.long)
            c = concs[:, sel, :].permute(1, 0, 2).contiguous().reshape(-1, A1)                 # [Kq*M, A1]
            v = torch.as_tensor(np.asarray(q_counts, dtype=np.float64), device=dev)[:, None, :].expand(-1, M, -1)
            v = v.contiguous().reshape(-1, A1)
            out = torch.empty(c.shape[0], dtype=torch.float64, device=dev)
            check(lib.bear_dm_logprob(ptr(c), c.shape[0], ptr(v), c.shape[0], A1, ptr(out), _lib.stream()))
            return out.reshape(-1, M).sum(0).cpu().numpy()
        return prob_func
    if get_map:
        log_probs = torch.log(concs / concs.sum(-1, keepdim=True))[None]                       # [1, M, K, A1]
    else:
        log_probs = log_gamma.log_gamma_device(concs, mc_samples, seed)                        # [mc, M, K, A1]
        check(lib.bear_log_normalize(ptr(log_probs), log_probs.numel() // A1, A1, _lib.stream()))
    arr = log_probs.permute(2, 3, 1, 0).contiguous().cpu().numpy()                              # [K, A1, M, mc]
    if output == 'numpy':
        return arr
    if output == 'df':
        import pandas as pd
        kp1mers = cross_str_arrays(kmers, core.alphabets_en[alphabet_name])
        df = pd.DataFrame(arr.reshape(K * A1, M * mc_samples), index=kp1mers)
        df.columns = np.arange(len(df.columns))
        return df
    return _PDF(kmers, arr, alphabet_name, summed)


def lookup_counts(data, kmers, alphabet_name):
    """Counts [Kq, num_ds, A1] (float64, device) of the query k-mer strings in a resident table;
    zeros for k-mers that are absent.  One packed-code join on the device against the table's cached
    sorted index (``KmerTable.sorted_index``); a k-mer that occurs in several rows (tables concatenated
    from several files) gets the sum of its rows, as a counter would give."""
    table = data.table
    codes, lag = dataloader.encode_kmers(kmers, alphabet_name)
    if lag != table.lag:
        raise ValueError('query k-mers have length %d, the table has lag %d' % (lag, table.lag))
    k, c = table.device_tensors()
    q = torch.from_numpy(codes.view(np.int64)).to(k.device)
    if table.num_rows == 0:
        found = torch.zeros_like(q, dtype=torch.bool)
        return torch.zeros((q.numel(), c.shape[0], c.shape[1]), dtype=torch.float64, device=k.device), found
    keys, order = table.sorted_index()
    lo = torch.searchsorted(keys, q)
    hi = torch.searchsorted(keys, q, right=True)
    mult = hi - lo                                           # rows holding the k-mer
    found = mult > 0
    last = table.num_rows - 1
    out = c[:, :, order[lo.clamp(max=last)]].permute(2, 0, 1).to(torch.float64) * found[:, None, None]
    for d in range(1, int(mult.max()) if q.numel() else 0):  # repeated keys: add the other rows
        more = mult > d
        out += c[:, :, order[(lo + d).clamp(max=last)]].permute(2, 0, 1).to(torch.float64) * more[:, None, None]
    return out.contiguous(), found


def parse_var(var):
    """'AAG23CC' -> ('AAG', 'CC', 23); insertions and deletions allowed (get_var_probs.py:336-341)."""
    digits = [i for i, ch in enumerate(var) if ch.isnumeric()]
    lo, hi = digits[0], digits[0] + len(digits)
    return var[:lo], var[hi:], int(var[lo:hi])


def _windows(vars_, wt_seq, lag):
    """(wt window, mutant window) around each variant in the padded wild type (get_var_probs.py:293-334)."""
    out = []
    for wt_aa, mt_aa, pos in vars_:
        pos = pos + lag
        assert wt_aa == wt_seq[pos:pos + len(wt_aa)]
        tail = wt_seq[pos + len(wt_aa):pos + lag + len(wt_aa)]
        out.append((wt_seq[pos - lag:pos + lag + len(wt_aa)], wt_seq[pos - lag:pos] + mt_aa + tail))
    return out


def _kp1mers(win, lag):
    return [win[i:i + lag + 1] for i in range(len(win) - lag)]


def _models_setup(bear_path, lag, alphabet_name, h, data, vans, kmc_path, reference_compatible=True):
    if kmc_path is not None:
        raise NotImplementedError('the KMC random-access counter needs py_kmc_api, which is outside this hot path')
    if bear_path is not None:
        lag, alphabet_name, h_bear, ar_func, data = load_bear(bear_path, reference_compatible)
        if h is None:
            h = np.array([h_bear])
        len_h = len(h)
    else:
        assert lag is not None and alphabet_name is not None and data is not None and len(vans) > 0
        len_h, ar_func = 0, None
    return lag, alphabet_name, h, ar_func, data, len_h


def get_bear_probs(bear_path, wt_seq, vars_, train_col, mc_samples=41, vans=[0.1, 1, 10], get_map=False,
                   lag=None, alphabet_name=None, h=None, data=None,
                   kmc_path=None, kmc_reverse=False, kmc_no_end=False, seed=None):
    """Posterior-predictive log-probability ratios of variants vs the wild type
    (get_var_probs.py:343-454).  Returns [num variants, num models, mc_samples] ([.., num models]
    with get_map); models = (AR if get_map) + BEAR h's + BMM vans."""
    lag, alphabet_name, h, ar_func, data, len_h = _models_setup(bear_path, lag, alphabet_name, h, data, vans, kmc_path)
    wt_seq = lag * '[' + wt_seq + ']'
    wins = _windows([parse_var(v) for v in vars_], wt_seq, lag)
    all_kmers = np.array(sorted({k[:-1] for w, m in wins for k in _kp1mers(w, lag) + _kp1mers(m, lag)}))
    if get_map:
        mc_samples = 1
    counts, _ = lookup_counts(data, all_kmers, alphabet_name)
    pdf = get_pdf(all_kmers, counts, h, ar_func, mc_samples, vans, train_col, alphabet_name, get_map, seed=seed)
    scores = np.stack([pdf(_kp1mers(m, lag)) - pdf(_kp1mers(w, lag)) for w, m in wins])
    return scores[..., 0] if get_map else scores


def get_bear_probs_seqs(bear_path, seqs, train_col, mc_samples=41, vans=[0.1, 1, 10], get_map=False, get_marg=False,
                        lag=None, alphabet_name=None, h=None, data=None,
                        kmc_path=None, kmc_reverse=False, no_ends=False, seed=None):
    """Posterior-predictive log-probabilities of whole sequences (get_var_probs.py:510-631).
    Returns [num sequences, num models, mc_samples] ([.., num models] with get_map / get_marg)."""
    lag, alphabet_name, h, ar_func, data, len_h = _models_setup(bear_path, lag, alphabet_name, h, data, vans, kmc_path)
    alphabet = core.alphabets_en[alphabet_name]
    if not no_ends:
        seqs = [lag * '[' + s + ']' for s in seqs]
    for s in seqs:
        assert len(s.replace('[', '').replace(']', '')) >= lag
    all_kmers = np.array(sorted({k[:-1] for s in seqs for k in _kp1mers(s, lag)}))
    if get_map or get_marg:
        mc_samples = 1
    counts, _ = lookup_counts(data, all_kmers, alphabet_name)
    pdf = get_pdf(all_kmers, counts, h, ar_func, mc_samples, vans, train_col, alphabet_name, get_map, get_marg, seed=seed)
    if get_marg:
        letters = {ch: b for b, ch in enumerate(alphabet)}
        rows = []
        for s in seqs:
            agg = {}
            for k in _kp1mers(s, lag):
                agg.setdefault(k[:-1], np.zeros(len(alphabet)))[letters[k[-1]]] += 1
            rows.append(pdf(list(agg), np.stack(list(agg.values()))))
        return np.stack(rows)
    scores = np.stack([pdf(_kp1mers(s, lag)) for s in seqs])
    return scores[..., 0] if get_map else scores
