"""Generates tests/golden/*.json|npz.

The reference itself cannot be imported in this image (tensorflow, tensorflow_probability and
tensorflow_io are absent), so the golden vectors are produced from
  (a) the SciPy known-answer formulas the reference's own tests assert against
      (bear_model/tests/test_core.py:23-26, test_dataloader.py:42-46, test_run.py:26-30) evaluated on
      the bundled table bear_model/data/ysd1_lag_5_file_0_preshuf.tsv, and
  (b) the oracle (oracle/bear_oracle.py), once it reproduces (a), for quantities the reference does not
      pin (gradients, head outputs) on a small seeded synthetic table.
Run from the repo root:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np
import torch
from scipy.special import loggamma

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import bear_oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
YSD1 = os.path.join(ROOT, 'bear_b200', 'data', 'ysd1_lag_5_file_0_preshuf.tsv')
EPS = 1e-7


def scipy_bmm(counts, alpha):
    """bear_model/tests/test_dataloader.py:42-46"""
    return np.sum((np.sum(loggamma(counts[:, :, None, :] + alpha[:, None]), axis=-1)
                   - loggamma(np.sum(counts[:, :, None, :] + alpha[:, None], axis=-1)))
                  - (np.sum(loggamma(0 * counts[:, :, None, :] + alpha[:, None]), axis=-1)
                     - loggamma(np.sum(0 * counts[:, :, None, :] + alpha[:, None], axis=-1))), axis=0)


def main():
    kmers, counts = O.read_tsv(YSD1, 3)
    alpha = np.array([0.1, 1.0, 10.0])
    out = {
        'first_batch_kmers': kmers[:3],
        'first_batch_counts': counts[:3].tolist(),
        'num_rows': len(kmers),
        'column_sums': counts.sum((0, 2)).tolist(),
        'bmm_likelihood': scipy_bmm(counts, alpha).tolist(),
        'bmm_likelihood_alpha_plus_eps_train': scipy_bmm(counts, alpha + EPS)[0].tolist(),
    }
    tot0 = counts[:, 0].sum()
    out['perp_van_train'] = np.exp(-np.array(out['bmm_likelihood_alpha_plus_eps_train']) / tot0).tolist()
    # heldout BMM (train column 0 -> test column 1): conc = train + van + eps, SciPy formula
    tr, te = counts[:, 0], counts[:, 1]
    ll, acc = [], []
    for v in alpha:
        conc = tr + v + EPS
        ll.append(float(np.sum(np.sum(loggamma(conc + te) - loggamma(conc), -1)
                               - (loggamma(conc.sum(-1) + te.sum(-1)) - loggamma(conc.sum(-1))))))
        acc.append(float(np.sum(te[np.arange(len(te)), np.argmax(conc, -1)]) / te.sum()))
    out['heldout_ll_van'] = ll
    out['heldout_perp_van'] = np.exp(-np.array(ll) / te.sum()).tolist()
    out['heldout_acc_van'] = acc
    with open(os.path.join(HERE, 'ysd1_known_answers.json'), 'w') as fh:
        json.dump(out, fh, indent=1)

    # small seeded synthetic table + oracle outputs (unpinned-by-reference quantities)
    rng = np.random.default_rng(20211012)
    K, lag = 257, 7
    codes = rng.integers(0, 4 ** lag, size=K, dtype=np.uint64)
    nstart = np.where(rng.random(K) < 0.1, rng.integers(1, lag + 1, size=K), 0).astype(np.uint64)
    for i in np.flatnonzero(nstart):
        codes[i] &= np.uint64((1 << (2 * (lag - int(nstart[i])))) - 1)
    codes |= nstart << np.uint64(58)
    tot = rng.poisson(3.0, size=(K, 2)) * (rng.random((K, 2)) < 0.85)
    p = rng.dirichlet(0.3 * np.ones(5), size=(K, 2))
    counts = np.stack([[rng.multinomial(tot[i, g], p[i, g]) for g in range(2)] for i in range(K)]).astype(np.int64)
    strs = []
    for c in codes:
        ns, v = int(c >> np.uint64(58)), int(c & np.uint64((1 << 58) - 1))
        strs.append(''.join('[' if j < ns else 'ACGT'[(v >> (2 * (lag - 1 - j))) & 3] for j in range(lag)))
    oh = O.one_hot(strs)
    gen = torch.Generator().manual_seed(7)
    mat = O.init_linear(lag, 4, gen)[0] * 10.0
    hs = torch.tensor(-1.25, dtype=torch.float64)
    res = {}
    for mode, train_ar in (('bear', False), ('ar', True)):
        loss, ll, grads = O.train_step_grads(oh, torch.tensor(counts[:, 0], dtype=torch.float64), hs, [mat], 'linear',
                                             1000, train_ar)
        res[mode + '_loss'] = float(loss)
        res[mode + '_ll'] = ll.numpy()
        res[mode + '_dh'] = float(grads[0])
        res[mode + '_dmat'] = grads[1].numpy()
    f = O.ar_linear(oh, [mat])
    ev = O.evaluation([(oh, f, counts[:, 1].astype(float), counts[:, 0].astype(float))], torch.tensor(0.3, dtype=torch.float64),
                      np.array([0.5, 2.0]))
    np.savez(os.path.join(HERE, 'synthetic_lag7.npz'), codes=codes, counts=counts, kmers=np.array(strs), mat=mat.numpy(),
             h_signed=float(hs), num_kmers=1000, f=f.numpy(), eval=np.concatenate([np.atleast_1d(x.numpy()) for x in ev]),
             **res)
    print('wrote golden fixtures to', HERE)


if __name__ == '__main__':
    main()
