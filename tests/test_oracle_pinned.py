"""Pins the CPU oracle (oracle/bear_oracle.py) against the reference's own known-answer tests, golden
vectors and published results.  The reference cannot be imported here (TensorFlow absent), so these
are the anchors SURVEY.md 8(c) lists; each test cites the reference test it restates."""
import json
import os

import numpy as np
import pytest
import torch
from scipy import stats as st
from scipy.special import loggamma

from conftest import GOLDEN, SPARSE, YSD1
from oracle import bear_oracle as O

EPS = 1e-7


@pytest.fixture(scope='module')
def ysd1():
    return O.read_tsv(YSD1, 3)


@pytest.fixture(scope='module')
def known():
    with open(os.path.join(GOLDEN, 'ysd1_known_answers.json')) as fh:
        return json.load(fh)


def test_dm_counts_log_prob_matches_scipy_formula():
    """bear_model/tests/test_core.py:7-26 (broadcast conc [5, A+1] against counts [3, 5, A+1])"""
    rng = np.random.default_rng(0)
    shape, A = np.array([3, 5]), 4
    trans = rng.poisson(size=np.r_[shape, A + 1]).astype(float)
    total = trans.sum(-1)
    conc = rng.exponential(size=np.r_[shape[1], A + 1])
    sum_conc = conc.sum(-1)
    want = (np.sum(loggamma(conc + trans) - loggamma(conc), -1) - (loggamma(sum_conc + total) - loggamma(sum_conc)))
    for cancel in (True, False):
        got = O.dm_counts_log_prob(total, conc, trans, with_cancelling_term=cancel).numpy()
        assert np.allclose(got, want, rtol=1e-12, atol=1e-12)
    assert np.all(O.ml_output_noiseless(np.broadcast_to(conc, trans.shape)).numpy() == np.tile(np.argmax(conc, -1)[None], [3, 1]))


def test_mn_counts_log_prob_matches_formula():
    """bear_model/tests/test_core.py:42-60"""
    rng = np.random.default_rng(1)
    trans = rng.poisson(size=(3, 5, 5)).astype(float)
    conc = rng.exponential(size=(5, 5))
    probs = conc / conc.sum(-1, keepdims=True)
    got = O.mn_counts_log_prob(trans.sum(-1), probs, trans).numpy()
    assert np.allclose(got, np.sum(np.log(probs) * trans, -1))
    # multiply_no_nan: zero counts ignore log(0)
    p0 = np.array([[0.0, 0.5, 0.5]])
    assert float(O.mn_counts_log_prob(np.array([2.0]), p0, np.array([[0.0, 1.0, 1.0]]))) == pytest.approx(2 * np.log(0.5))


def test_golden_first_batch(ysd1, known):
    """bear_model/tests/test_dataloader.py:20-32"""
    kmers, counts = ysd1
    assert kmers[:3] == ['TAATC', 'CGGTC', 'ACGCT'] == known['first_batch_kmers']
    want = [[[14837, 15127, 22260, 16279, 446], [5029, 5095, 7408, 5487, 134], [16, 16, 23, 17, 0]],
            [[61890, 729, 39733, 35956, 1017], [20524, 239, 13199, 12046, 309], [69, 0, 45, 39, 0]],
            [[13965, 23135, 73870, 37045, 1035], [4705, 7591, 24532, 12305, 385], [14, 25, 81, 39, 0]]]
    assert np.all(counts[:3] == np.array(want))
    assert counts.dtype == np.float64 and len(kmers) == 1365 and 1365 / 3 == 455
    assert sum('[' in k for k in kmers) == 341
    assert counts.sum((0, 2)).tolist() == [114584236.0, 38118742.0, 117834.0] and counts.max() == 254715


def test_bmm_likelihood_known_answer(ysd1, known):
    """bear_model/tests/test_dataloader.py:34-49; values from BASELINE.md"""
    _, counts = ysd1
    got = O.bmm_likelihood(counts, np.array([0.1, 1.0, 10.0])).numpy()
    want = np.array([[-1.5271257134588018e8, -1.5270905139595887e8, -1.5274538628200924e8],
                     [-5.0819958884193577e7, -5.0816105408838786e7, -5.0848897074304186e7],
                     [-1.6336475286893066e5, -1.6157309667778228e5, -1.7003849692050868e5]])
    assert np.allclose(got, want, rtol=1e-12)
    assert np.allclose(got, np.array(known['bmm_likelihood']), rtol=1e-12)


def test_evaluation_ties_to_bmm_closed_form(ysd1, known):
    """bear_model/tests/test_run.py:26-30: evaluation's ll_van / perp_van on the train column ==
    bmm_likelihood(alpha + eps)[0]; plus the heldout numbers of BASELINE.md / docs/usage.rst:255-265."""
    kmers, counts = ysd1
    oh = O.one_hot(kmers)
    gen = torch.Generator().manual_seed(0)
    f = O.ar_linear(oh, O.init_linear(5, 4, gen))
    van = np.array([0.1, 1.0, 10.0])
    out = O.evaluation([(oh, f, counts[:, 0], None)], torch.tensor(1.0, dtype=torch.float64), van)
    train_liks = O.bmm_likelihood(counts, van + EPS)[0].numpy()
    assert np.allclose(out[2].numpy(), train_liks, rtol=1e-13)
    assert np.allclose(out[2].numpy(), [-152712571.34208855, -152709051.39618367, -152745386.2824309], rtol=1e-12)
    assert np.allclose(out[5].numpy(), np.exp(-train_liks / counts[:, 0].sum()))
    assert np.allclose(out[5].numpy(), known['perp_van_train'], rtol=1e-12)
    held = O.evaluation([(oh, f, counts[:, 1], counts[:, 0])], torch.tensor(1.0, dtype=torch.float64), van)
    assert np.allclose(held[2].numpy(), known['heldout_ll_van'], rtol=1e-12)
    assert np.allclose(held[5].numpy(), [3.790636628, 3.790645212, 3.790733706], rtol=1e-9)
    assert np.allclose(held[8].numpy(), 0.3676534498, rtol=1e-9)
    # the published table rounds these to "BMM 3.79, 36.8 %" (docs/usage.rst:258)
    assert round(float(held[5][1]), 2) == 3.79 and round(100 * float(held[8][1]), 1) == 36.8


def test_one_hot_symbol_order():
    """core.py:142-174: A,C,G,T,'[' columns; unknown symbol -> zero row"""
    oh = O.one_hot(['AC[', 'GTN'])
    assert oh.shape == (2, 3, 5)
    assert oh[0].tolist() == [[1, 0, 0, 0, 0], [0, 1, 0, 0, 0], [0, 0, 0, 0, 1]]
    assert oh[1].tolist() == [[0, 0, 1, 0, 0], [0, 0, 0, 1, 0], [0, 0, 0, 0, 0]]
    assert O.one_hot([b'AR['], 'prot').shape == (1, 3, 21)


def test_sparse_reader_matches_dense_twin():
    """dataloader.py:52-109 on data/ex_seqs_kmap_for_var_pred.csv vs its TSV twin data/kmaps/ex_seqs_lag_3_file_0.tsv"""
    ks, cs = O.read_sparse(SPARSE, 1)
    kd, cd = O.read_tsv(os.path.join(os.path.dirname(SPARSE), 'kmaps', 'ex_seqs_lag_3_file_0.tsv'), 1)
    assert dict(zip(ks, map(lambda x: x.tolist(), cs))) == dict(zip(kd, map(lambda x: x.tolist(), cd)))
    assert cs[ks.index('TTT'), 0].tolist() == [1, 0, 0, 4, 2]       # tests/test_var_prob.py:14


def test_map_scores_known_answer():
    """bear_model/tests/test_var_prob.py:60-78: MAP variant scores of the toy sequences equal
    log((seen + van) / (all + (A+1) van)) sums."""
    ks, cs = O.read_sparse(SPARSE, 1)
    vans = np.array([0.1, 1, 10])
    concs = O.get_pdf_concs(cs[:, 0, :], None, None, vans, get_map=True)
    lp = O.get_pdf_map(concs)                                        # [V, K, A+1]
    letters = 'ACGT]'

    def tp(kmer, b):
        return lp[:, ks.index(kmer), letters.index(b)]

    def quot(seen, all_, van):
        return np.log((seen + van) / (all_ + 5 * van))
    # wt TTTAT -> variant A3T: mutant window transitions minus wild-type ones (test_var_prob.py:44-45,70-71);
    # k-mers absent from the table (ATT, ATA...) fall back to the prior, here only seen ones are compared
    assert np.allclose(tp('TTT', 'T'), quot(4, 7, vans))
    assert np.allclose(tp('TTT', ']'), quot(2, 7, vans))
    assert np.allclose(tp('TTT', 'A'), quot(1, 7, vans))
    assert np.allclose(tp('TTA', 'T'), quot(1, 1, vans))
    assert np.allclose(tp('[TT', 'T'), quot(3, 4, vans))
    assert np.allclose(tp('[TT', 'C'), quot(1, 4, vans))


def test_log_gamma_sampler_ks():
    """bear_model/tests/test_log_gamma.py:9-19 (smaller n)"""
    rng = np.random.default_rng(0)
    concs = np.array([0.01, 0.1, 0.5, 0.99, 1, 5, 100])
    n = 20000
    tile = (np.ones([len(concs), n]) * concs[:, None]).flatten()
    samples = O.log_gamma_sample(tile, [2], rng).reshape(2, len(concs), n)
    for i, c in enumerate(concs):
        assert st.kstest(np.exp(samples[:, i].flatten()), cdf='gamma', args=[c]).pvalue > 0.1 / 6


def test_gradients_match_finite_differences(ysd1):
    """The reference pins no gradients; the oracle's autograd gradients are checked by central differences."""
    kmers, counts = ysd1
    oh = O.one_hot(kmers[:200])
    c = torch.tensor(counts[:200, 0])
    gen = torch.Generator().manual_seed(3)
    mat = O.init_linear(5, 4, gen)[0] * 5
    hs = torch.tensor(-0.4, dtype=torch.float64)
    for train_ar in (False, True):
        loss, _, grads = O.train_step_grads(oh, c, hs, [mat], 'linear', 1365, train_ar)
        d = 1e-4
        lp, _ = O.train_loss(oh, c, hs + d, lambda x: O.ar_linear(x, [mat]), 1365, train_ar)
        lm, _ = O.train_loss(oh, c, hs - d, lambda x: O.ar_linear(x, [mat]), 1365, train_ar)
        fd = float(lp - lm) / (2 * d)
        assert abs(fd - float(grads[0])) <= 1e-5 * max(abs(fd), 1.0)
        for idx in [(0, 0, 0), (2, 3, 1), (4, 1, 4)]:
            mp_, mm_ = mat.clone(), mat.clone()
            mp_[idx] += d
            mm_[idx] -= d
            lp, _ = O.train_loss(oh, c, hs, lambda x: O.ar_linear(x, [mp_]), 1365, train_ar)
            lm, _ = O.train_loss(oh, c, hs, lambda x: O.ar_linear(x, [mm_]), 1365, train_ar)
            fd = float(lp - lm) / (2 * d)
            assert abs(fd - float(grads[1][idx])) <= 1e-5 * max(abs(fd), 1.0)


def test_keras_adam_restatement():
    """tf.keras.optimizers.Adam (OptimizerV2): theta -= lr*sqrt(1-b2^t)/(1-b1^t) * m/(sqrt(v)+1e-7)"""
    p = [torch.tensor([1.0, -2.0], dtype=torch.float64)]
    opt = O.KerasAdam(p, 0.1)
    g = torch.tensor([0.5, -4.0], dtype=torch.float64)
    opt.apply(p, [g])
    m, v = 0.1 * g, 0.001 * g * g
    lr_t = 0.1 * np.sqrt(1 - 0.999) / (1 - 0.9)
    want = torch.tensor([1.0, -2.0], dtype=torch.float64) - lr_t * m / (v.sqrt() + 1e-7)
    assert torch.allclose(p[0], want, rtol=1e-15)
    assert abs(float(p[0][0]) - (1.0 - 0.1)) < 1e-6       # first Adam step moves by ~lr


def test_cancelling_term_is_rounding_noise(ysd1):
    """core.py:74 adds and subtracts log_combinations; both forms agree to ~1e-15 relative per k-mer."""
    _, counts = ysd1
    c = counts[:, 0]
    conc = np.random.default_rng(0).dirichlet(np.ones(5), size=len(c)) / 0.0433 + EPS
    a = O.dm_counts_log_prob(c.sum(-1), conc, c, True).numpy()
    b = O.dm_counts_log_prob(c.sum(-1), conc, c, False).numpy()
    assert np.max(np.abs(a - b) / np.abs(b)) < 1e-13


def test_vectorised_one_hot_matches_the_per_character_restatement():
    """bench.py's CPU arm builds the one-hot input from byte strings every step (as the reference's tf.data map does,
    bear_net.py:268-273) with the vectorised form of core.tf_one_hot; it must equal the per-character restatement."""
    rng = np.random.default_rng(0)
    kmers = np.array([''.join(rng.choice(list('ACGT['), size=7)).encode() for _ in range(200)] + [b'ACGTNNN'], dtype='S7')
    assert torch.equal(O.one_hot_bytes(kmers), O.one_hot(list(kmers)))
