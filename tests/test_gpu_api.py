"""GPU tests of the remaining reference-facing API: bear_ref, the posterior sampler, get_var_probs and the
config-driven scripts -- each against the oracle or the reference tests' known answers."""
import configparser
import os

import numpy as np
import pytest
import torch
from scipy import stats as st
from scipy.special import logsumexp

from conftest import ROOT, SPARSE, YSD1

pytestmark = pytest.mark.gpu


def _oracle():
    from oracle import bear_oracle as O
    return O


def rel_err(got, want):
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    return np.max(np.abs(got - want)) / max(np.max(np.abs(want)), 1e-300)


# ------------------------------------------------------------------------------------------------
# bear_ref
# ------------------------------------------------------------------------------------------------
def _oracle_ref_loss(O, oh, counts, ref, hs, tau_s, nw_s, net_params, net, K, train_ar):
    hs, tau_s, nw_s = [x.clone().requires_grad_(True) for x in (hs, tau_s, nw_s)]
    ps = [p.clone().requires_grad_(True) for p in net_params]
    refc = O.ref_counts_map(ref, 4)
    net_func = (lambda x: O.ar_stop(x, 4)) if net == 'stop' else (lambda x: O.AR_FUNCS[net](x, ps))
    loss, ll = O.train_loss(oh, counts, hs, lambda x: O.ar_ref(x, refc, tau_s, nw_s, net_func, 4), K, train_ar)
    leaves = [hs, tau_s, nw_s] + ps
    grads = torch.autograd.grad(loss, leaves, allow_unused=True)
    return loss.detach(), [torch.zeros_like(l) if g is None else g for g, l in zip(grads, leaves)]


@pytest.mark.parametrize('net,train_ar', [('stop', False), ('stop', True), ('linear', False)])
def test_bear_ref_train_step_matches_oracle(cuda, net, train_ar):
    """bear_ref.train (bear_ref.py:207-259, 262-389): one SGD step recovers the oracle's gradients for
    [h_signed, tau_signed, net_weight_signed, *net params]; config bear_stop_bear.cfg / bear_stop_ar.cfg."""
    from bear_b200 import ar_funcs, bear_ref, dataloader as dl
    O = _oracle()
    data = dl.dataloader(YSD1, 'dna', 1500, 3)
    K = data.table.num_rows
    make = getattr(ar_funcs, 'make_ar_func_' + net)
    torch.manual_seed(11)
    p0, _, _ = bear_ref._create_params(5, 4, make, {})
    p0 = [p.clone() for p in p0]
    p0[0].fill_(-0.3)
    ls = []
    params, h_signed, ar_func = bear_ref.train(data, K, 1, 0, 2, 'dna', 5, make, {}, 1e-3, 'SGD', train_ar,
                                               params_restart=p0, loss_save=ls)
    kmers, counts = O.read_tsv(YSD1, 3)
    loss, grads = _oracle_ref_loss(O, O.one_hot(kmers), torch.tensor(counts[:, 0]), counts[:, 2],
                                   p0[0].cpu(), p0[1].cpu(), p0[2].cpu(), [p.cpu() for p in p0[3:]], net, K, train_ar)
    assert abs(-ls[0] - float(loss)) <= 1e-10 * abs(float(loss))
    for new, old, g in zip(params, p0, grads):
        got = (old.cpu() - new.cpu()) / 1e-3
        assert np.max(np.abs(got.numpy() - g.numpy())) <= 1e-7 * max(float(g.abs().max()), 1e-6 * abs(float(loss)))


def test_bear_ref_evaluation_matches_oracle(cuda):
    from bear_b200 import ar_funcs, bear_ref, dataloader as dl
    O = _oracle()
    data = dl.dataloader(YSD1, 'dna', 400, 3)
    params, h_signed, ar_func = bear_ref._create_params(5, 4, ar_funcs.make_ar_func_stop, {})
    van = np.array([0.1, 1.0, 10.0])
    kmers, counts = O.read_tsv(YSD1, 3)
    oh = O.one_hot(kmers)
    refc = O.ref_counts_map(counts[:, 2], 4)
    f = O.ar_ref(oh, refc, params[1].cpu(), params[2].cpu(), lambda x: O.ar_stop(x, 4), 4)
    for train_col, test_col in ((-1, 0), (0, 1)):
        got = bear_ref.evaluation(data, train_col, test_col, 2, 'dna', 0.0142, ar_func, van, seed=-1)
        want = O.evaluation([(oh, f, counts[:, test_col], counts[:, train_col] if train_col >= 0 else None)],
                            torch.tensor(0.0142, dtype=torch.float64), van)
        for g, w in zip(got, want):
            assert rel_err(g.numpy(), w.numpy()) <= 1e-10
    # reference_compatible: upstream conditions on the mapped reference column (bear_ref.py:397) -- both modes are pinned
    got = bear_ref.evaluation(data, 0, 1, 2, 'dna', 0.0142, ar_func, van, seed=-1, reference_compatible=True)
    want = O.evaluation([(oh, f, counts[:, 1], refc)], torch.tensor(0.0142, dtype=torch.float64), van)
    for g, w in zip(got, want):
        assert rel_err(g.numpy(), w.numpy()) <= 1e-10
    # train_test path ties to the BMM closed form, as tests/test_run.py:47-51 asserts for the ref script
    assert np.allclose(bear_ref.evaluation(data, -1, 0, 2, 'dna', 1.0, ar_func, van, seed=-1)[2].numpy(),
                       [-152712571.34208855, -152709051.39618367, -152745386.2824309], rtol=1e-11)


@pytest.mark.parametrize('lag,train_col', [(7, -1), (13, 0)])
def test_bear_ref_evaluation_linear_net_fused_matches_oracle(cuda, lag, train_col):
    """bear_ref.evaluation with a linear embedded net runs in the fused kernel (bear_ref_eval_step, reference head in
    registers, bear_ref.py:63-68 + 391-437); checked on a synthetic table with start-padded k-mers, ragged batches and a
    sparse reference column."""
    from bear_b200 import ar_funcs, bear_ref, dataloader as dl
    from test_gpu_parity import synth_table
    O = _oracle()
    K = 2500
    codes, counts = synth_table(K, lag, 3, seed=50 + lag, start_frac=0.2)
    counts[:, 2] = np.random.default_rng(lag).poisson(0.4, size=(K, 5))
    data = dl.KmerDataset(dl.KmerTable.from_arrays((codes, lag), counts, 'dna'), 777)
    torch.manual_seed(lag)
    params, h_signed, ar_func = bear_ref._create_params(lag, 4, ar_funcs.make_ar_func_linear, {})
    params[3].mul_(30.0)
    params[2].fill_(0.3)                           # net weight e^0.3: the net matters next to the reference counts
    van = np.array([0.5, 2.0])
    oh = O.one_hot(dl.decode_kmers(codes, lag, 'dna'))
    refc = O.ref_counts_map(counts[:, 2].astype(np.float64), 4)
    f = O.ar_ref(oh, refc, params[1].cpu(), params[2].cpu(), lambda x: O.ar_linear(x, [params[3].cpu()]), 4)
    test_col = 1
    got = bear_ref.evaluation(data, train_col, test_col, 2, 'dna', 0.37, ar_func, van, seed=-1)
    want = O.evaluation([(oh, f, counts[:, test_col].astype(np.float64),
                          counts[:, train_col].astype(np.float64) if train_col >= 0 else None)],
                        torch.tensor(0.37, dtype=torch.float64), van)
    for g, w in zip(got, want):
        assert rel_err(g.numpy(), w.numpy()) <= 1e-10


# ------------------------------------------------------------------------------------------------
# sampler + get_var_probs
# ------------------------------------------------------------------------------------------------
def test_log_gamma_sampler_ks(cuda):
    """reference tests/test_log_gamma.py:9-19"""
    from bear_b200 import log_gamma
    concs = np.array([0.01, 0.1, 0.5, 0.99, 1, 5, 100])
    n, n_tile = 100000, 3
    tile = (np.ones([len(concs), n]) * concs[:, None]).flatten()
    samples = log_gamma.log_gamma(tile, size=[n_tile], seed=1).reshape([n_tile, len(concs), n])
    assert np.all(np.isfinite(samples))
    for i, c in enumerate(concs):
        assert st.kstest(np.exp(samples[:, i].flatten()), cdf='gamma', args=[c]).pvalue > 0.1 / 6
    # tiny concentrations stay finite in log space where Gamma samples underflow to 0
    tiny = log_gamma.log_gamma(np.full(1000, 1e-7), size=[2], seed=2)
    assert np.all(np.isfinite(tiny)) and tiny.shape == (2, 1000) and np.median(tiny) < -1e5
    # same seed -> same draws, different shape of the launch or not
    assert np.array_equal(log_gamma.log_gamma(concs, size=[4], seed=3), log_gamma.log_gamma(concs, size=[4], seed=3))


def _toy_data():
    from bear_b200 import dataloader
    return dataloader.sparse_dataloader(SPARSE, 'dna', 500, 1)


def test_get_bear_probs_map_known_answer(cuda):
    """reference tests/test_var_prob.py:60-78 (MAP scores exact)"""
    from bear_b200 import get_var_probs
    vans = np.array([0.1, 1, 10])
    scores = get_var_probs.get_bear_probs(None, 'TTTAT', np.array(['A3T', 'T2C']), 0, data=_toy_data(), get_map=True,
                                          vans=vans, lag=3, alphabet_name='dna')

    def q(seen, all_, van):
        return np.log((seen + van) / (all_ + 5 * van))
    true = np.empty([2, 3])
    for i, van in enumerate(vans):
        true[0, i] = (2 * q(4, 7, van) + 1 * q(2, 7, van)) - (1 * q(1, 7, van) + 2 * q(1, 1, van))
        true[1, i] = (q(1, 4, van) + q(0, 1, van) + 2 * q(0, 0, van)) - (q(3, 4, van) + q(1, 7, van) + 2 * q(1, 1, van))
    assert np.allclose(scores, true)


def test_get_bear_probs_mc_statistical(cuda):
    """reference tests/test_var_prob.py:20-58 (MC mean within 2 % of Beta-sample truth)"""
    from bear_b200 import get_var_probs
    vans = np.array([0.1, 1, 10])
    scores = get_var_probs.get_bear_probs(None, 'TTTAT', np.array(['A3T', 'T2C']), 0, data=_toy_data(),
                                          mc_samples=500000, vans=vans, lag=3, alphabet_name='dna', seed=5)
    assert scores.shape == (2, 3, 500000)
    rng = np.random.default_rng(0)

    def d(seen, all_, van, n=500000):
        return np.average(np.log(st.beta.rvs(seen + van, all_ - seen + 4 * van, size=n, random_state=rng)))
    true = np.empty([2, 3])
    for i, van in enumerate(vans):
        true[0, i] = (2 * d(4, 7, van) + 1 * d(2, 7, van)) - (1 * d(1, 7, van) + 2 * d(1, 1, van))
        true[1, i] = (d(1, 4, van) + d(0, 1, van) + 2 * d(0, 0, van)) - (d(3, 4, van) + d(1, 7, van) + 2 * d(1, 1, van))
    assert np.all(np.abs((np.average(scores, axis=-1) - true) / true) < 0.02)


def test_get_bear_probs_seqs(cuda):
    """reference tests/test_var_prob.py:81-173: MAP exact, marginal and MC within 1 %"""
    from bear_b200 import get_var_probs
    seqs = ['TTTAT', 'TTCAT', 'TTTTTTTTTT']
    vans = np.array([0.1, 1, 10])
    kw = dict(vans=vans, lag=3, alphabet_name='dna')
    scores = get_var_probs.get_bear_probs_seqs(None, seqs, 0, data=_toy_data(), get_map=True, **kw)

    def q(seen, all_, van):
        return np.log((seen + van) / (all_ + 5 * van))
    true = np.empty([3, 3])
    for i, van in enumerate(vans):
        true[0, i] = 2 * q(4, 4, van) + q(3, 4, van) + q(1, 7, van) + 2 * q(1, 1, van)
        true[1, i] = 2 * q(4, 4, van) + q(1, 4, van) + q(0, 1, van) + 2 * q(0, 0, van)
        true[2, i] = 2 * q(4, 4, van) + q(3, 4, van) + 7 * q(4, 7, van) + q(2, 7, van)
    assert np.allclose(scores, true)

    N = 50000
    rng = np.random.default_rng(1)

    def d(seen, all_, van):
        return np.log(st.beta.rvs(seen + van, all_ - seen + 4 * van, size=N, random_state=rng))
    ts = np.empty([3, 3, N])
    for i, van in enumerate(vans):
        ts[0, i] = d(4, 4, van) + d(4, 4, van) + d(3, 4, van) + d(1, 7, van) + d(1, 1, van) + d(1, 1, van)
        ts[1, i] = d(4, 4, van) + d(4, 4, van) + d(1, 4, van) + d(0, 1, van) + d(0, 0, van) + d(0, 0, van)
        ttt = np.log(st.beta.rvs(4 + van, 2 + van, size=N, random_state=rng))
        mod = np.log(st.beta.rvs(6 + 2 * van, 1 + 3 * van, size=N, random_state=rng))
        ts[2, i] = d(4, 4, van) + d(4, 4, van) + d(3, 4, van) + 7 * (ttt + mod) + (np.log1p(-np.exp(ttt)) + mod)
    mc = get_var_probs.get_bear_probs_seqs(None, seqs, 0, data=_toy_data(), mc_samples=20000, seed=9, **kw)
    av = np.average(ts, axis=-1)
    assert np.all(np.abs((np.average(mc, axis=-1) - av) / av) < 0.01)
    margs = get_var_probs.get_bear_probs_seqs(None, seqs, 0, data=_toy_data(), get_marg=True, **kw)
    av = logsumexp(ts, axis=-1) - np.log(N)
    assert np.all(np.abs((margs - av) / av) < 0.01)


def test_get_pdf_with_bear_model_matches_oracle(cuda):
    """get_pdf concentrations / MAP output with an AR head and several h (get_var_probs.py:132-153,172)"""
    from bear_b200 import ar_funcs, get_var_probs
    O = _oracle()
    kmers, counts = O.read_tsv(YSD1, 3)
    kmers, counts = kmers[:64], counts[:64]
    torch.manual_seed(2)
    ar_func, params = ar_funcs.make_ar_func_linear(5, 4)
    hs, vans = np.array([0.05, 2.0]), [0.1, 1.0]
    got = get_var_probs.get_pdf(np.array(kmers), counts, hs, ar_func, 1, vans, 0, 'dna', True, output='numpy')
    ar_vals = O.ar_linear(O.one_hot(kmers), [params[0].cpu()]).numpy()
    want = O.get_pdf_map(O.get_pdf_concs(counts[:, 0, :], ar_vals, hs, vans, True))        # [M, K, A1]
    assert got.shape == (64, 5, 5, 1)
    assert rel_err(got[..., 0], np.transpose(want, (1, 2, 0))) <= 1e-12
    df = get_var_probs.get_pdf(np.array(kmers), counts, hs, ar_func, 1, vans, 0, 'dna', True, output='df')
    assert df.shape == (64 * 5, 5) and df.index[4] == kmers[0] + ']'
    assert np.allclose(df.loc[kmers[3] + 'G'].to_numpy(), got[3, 2, :, 0])


# ------------------------------------------------------------------------------------------------
# scripts / model-file layout
# ------------------------------------------------------------------------------------------------
def _config(name, out_dir, **over):
    config = configparser.ConfigParser()
    config.read(os.path.join(ROOT, 'bear_b200', 'models', 'config_files', name))
    config['general']['out_folder'] = str(out_dir) + '*'
    for k, v in over.items():
        sec, key = k.split('__')
        config[sec][key] = v
    return config


def _bmm_train_liks():
    from bear_b200 import dataloader
    data = dataloader.dataloader(YSD1, 'dna', 2000, 3)
    liks = dataloader.bmm_likelihood(data, np.array([0.1, 1., 10.]) + 1e-7)[0].numpy()
    tot = 114584236.0
    return liks, np.exp(-liks / tot)


def test_run_net_script(cuda, tmp_path):
    """reference tests/test_run.py:12-30 (bear_test.cfg: 1 epoch, linear head, train_ar=True)"""
    import dill
    from bear_b200 import ar_funcs, bear_net
    from bear_b200.models import train_bear_net
    out = tmp_path / 'net'
    exit_, ll_van, perp_van = train_bear_net.main(_config('bear_test.cfg', out))
    assert exit_ == 1
    liks, perp = _bmm_train_liks()
    assert np.allclose(liks, ll_van) and np.allclose(perp, perp_van)
    # model-file layout: config.cfg with [results], results.pickle = {'params': [h_signed, mat]}
    cfg = configparser.ConfigParser()
    cfg.read(out / 'config.cfg')
    for key in ('h', 'heldout_perplex_bear', 'heldout_perplex_ar', 'heldout_perplex_bmm', 'perplex_bmm',
                'heldout_accuracy_bear', 'loglikelihood_bear', 'file', 'out_folder'):
        assert key in cfg['results'], key
    with open(out / 'results.pickle', 'rb') as fh:
        params = dill.load(fh)['params']
    assert len(params) == 2 and params[0].shape == () and params[1].shape == (5, 5, 5)
    # restart from the saved folder, evaluate only (train = False, restart = True)
    out2 = tmp_path / 'net2'
    r = train_bear_net.main(_config('bear_test.cfg', out2, train__train='False', train__restart='True',
                                    train__restart_path=str(out)))
    assert r[0] == 1 and np.allclose(r[1], ll_van)
    p2, h2, f2 = bear_net.change_scope_params(5, 4, ar_funcs.make_ar_func_linear, {}, params)
    assert np.array_equal(p2[1].cpu().numpy(), params[1])
    # precision = float32 configs (models/train_bear_net.py:43) are accepted and computed in float64
    r32 = train_bear_net.main(_config('bear_test.cfg', tmp_path / 'net32', general__precision='float32'))
    assert r32[0] == 1 and np.allclose(r32[1], ll_van, rtol=1e-6)


def test_run_ref_script(cuda, tmp_path):
    """reference tests/test_run.py:32-51"""
    from bear_b200.models import train_bear_ref
    out = tmp_path / 'ref'
    exit_, ll_van, perp_van = train_bear_ref.main(_config('bear_test.cfg', out))
    assert exit_ == 1
    liks, perp = _bmm_train_liks()
    assert np.allclose(liks, ll_van) and np.allclose(perp, perp_van)
    cfg = configparser.ConfigParser()
    cfg.read(out / 'config.cfg')
    assert 'error_rate' in cfg['results'] and 'stop_rate' in cfg['results']


def test_lin_bear_config_converges_towards_published_table(cuda, tmp_path):
    """docs/usage.rst:255-265: Linear BEAR reaches heldout perplexity 3.79 / accuracy 36.8 % (h = 0.0433 after
    10000 steps); 600 steps are enough to be at the BMM-level perplexity and for h to head below 1."""
    from bear_b200.models import train_bear_net
    out = tmp_path / 'lin'
    train_bear_net.main(_config('bear_lin_bear.cfg', out, data__files_path='TEST', train__epochs='600'))
    cfg = configparser.ConfigParser()
    cfg.read(out / 'config.cfg')
    r = cfg['results']
    assert 3.78 < float(r['heldout_perplex_bear']) < 3.81
    assert 0.36 < float(r['heldout_accuracy_bear']) < 0.375
    assert float(r['h']) < 0.5
    assert float(r['heldout_perplex_ar']) > float(r['heldout_perplex_bear'])


def test_packed_shard_cache_trains_identically(cuda, tmp_path):
    """dataloader.pack_files / load_packed: a BEARPACK shard gives the same batches and the same training
    result as parsing the TSV."""
    from bear_b200 import ar_funcs, bear_net, dataloader
    out = str(tmp_path / 'ysd1.bearpack')
    dataloader.pack_files([YSD1], out, 'dna', 3)
    a = dataloader.dataloader(YSD1, 'dna', 400, 3)
    b = dataloader.load_packed(out, 400)
    ka, ca = next(iter(a))
    kb, cb = next(iter(b))
    assert np.array_equal(ka.numpy(), kb.numpy()) and torch.equal(ca, cb)
    torch.manual_seed(3)
    p0, _, _ = bear_net._create_params(5, 4, ar_funcs.make_ar_func_linear, {})
    p0 = [p.clone() for p in p0]
    ra = bear_net.train(a, 1365, 1, 0, 'dna', 5, ar_funcs.make_ar_func_linear, {}, 0.01, 'Adam', False, params_restart=p0)
    rb = bear_net.train(b, 1365, 1, 0, 'dna', 5, ar_funcs.make_ar_func_linear, {}, 0.01, 'Adam', False, params_restart=p0)
    assert torch.allclose(ra[0][1], rb[0][1], rtol=1e-12, atol=0) and float(ra[1]) == pytest.approx(float(rb[1]), rel=1e-12)


@pytest.mark.parametrize('alphabet,lag,n', [('dna', 20, 100003), ('dna', 3, 5), ('prot', 7, 4097)])
def test_upload_through_compact_format_is_bit_exact(cuda, alphabet, lag, n):
    """KmerTable.device_tensors() crosses the bus as byte planes + escapes and is expanded on the device
    (bear_expand_table): the resident table equals the host arrays bit for bit, padding rows are zero; a chunk can
    also be expanded at an unaligned destination row."""
    import numpy as np
    import torch
    from bear_b200 import _lib, dataloader as dl
    from bear_b200._lib import lib, check, ptr
    rng = np.random.default_rng(n)
    A1 = (20 if alphabet == 'prot' else 4) + 1
    if alphabet == 'prot':
        codes = rng.integers(0, 1 << (5 * lag), size=n, dtype=np.uint64)
    else:
        codes = rng.integers(0, 4 ** lag, size=n, dtype=np.uint64)
        ns = np.where(rng.random(n) < 0.2, rng.integers(1, lag + 1, size=n), 0).astype(np.uint64)
        for i in np.flatnonzero(ns):
            codes[i] &= np.uint64((1 << (2 * (lag - int(ns[i])))) - 1)
        codes |= ns << np.uint64(58)
    counts = rng.integers(0, 3, size=(n, 3, A1)) * rng.choice([1, 127, 255, 256, 4000000000 // 2], size=(n, 3, A1))
    table = dl.KmerTable.from_arrays((codes, lag), counts, alphabet)
    k, c = table.device_tensors()
    assert np.array_equal(k.cpu().numpy().view(np.uint64), table.kmers_host)
    assert np.array_equal(c.cpu().numpy().view(np.uint32), table.counts_host)
    # expansion at an unaligned destination (scalar stores)
    m = min(n, 1001)
    # 4-bit count planes: every count >= 15 travels as an escape; | 16: so do the start-run lengths (DNA / RNA)
    for bits in (8, 4, 12) + (() if alphabet == 'prot' else (8 | 16, 4 | 16, 12 | 16)):
        buf, esc, got_bits = table.compact_chunk(n - m, m, wire=bits)
        assert got_bits == bits and buf.numel() == table.compact_bytes(m, bits)
        k2 = torch.zeros(m + 8, dtype=torch.int64, device=cuda)
        c2 = torch.zeros((3, A1, m + 8), dtype=torch.int32, device=cuda)
        dbuf, desc = buf.to(cuda), (esc.to(cuda) if esc.numel() else None)
        check(lib.bear_expand_table(ptr(dbuf), ptr(desc), esc.shape[0], m, lag, _lib.ALPHABET_IDS[alphabet], 3, bits,
                                    ptr(k2), ptr(c2), m + 8, 3, _lib.stream()))
        assert np.array_equal(k2[3:3 + m].cpu().numpy().view(np.uint64), table.kmers_host[n - m:n])
        assert np.array_equal(c2[:, :, 3:3 + m].cpu().numpy().view(np.uint32), table.counts_host[:, :, n - m:n])
        assert int(c2[:, :, :3].abs().sum()) == 0 and int(c2[:, :, 3 + m:].abs().sum()) == 0
    # sparse counts (the benchmark regime) choose the 12-bit rank coding (protein, 21 letters of Poisson(0.6): the 4-bit
    # planes), and the upload through them is bit exact
    small = rng.poisson(0.6, size=(n, 3, A1)) * (rng.random((n, 3, A1)) < 0.999) + 40 * (rng.random((n, 3, A1)) < 0.001)
    sparse = dl.KmerTable.from_arrays((codes, lag), small, alphabet)
    assert sparse.compact_chunk(0, n)[2] & 15 == (4 if alphabet == 'prot' else 12) and table.compact_chunk(0, n)[2] & 15 == 8
    if alphabet == 'dna' and lag == 20:     # 1 % start-padded rows: the start-run lengths travel as escapes, 5 k-mer planes
        few = codes.copy()
        few[100:] &= np.uint64((1 << 58) - 1)
        t2 = dl.KmerTable.from_arrays((few, lag), small, alphabet)
        assert t2.compact_chunk(0, n)[2] == 12 | 16
        k4, c4 = t2.device_tensors()
        assert np.array_equal(k4.cpu().numpy().view(np.uint64), t2.kmers_host)
        assert np.array_equal(c4.cpu().numpy().view(np.uint32), t2.counts_host)
    k3, c3 = sparse.device_tensors()
    assert np.array_equal(k3.cpu().numpy().view(np.uint64), sparse.kmers_host)
    assert np.array_equal(c3.cpu().numpy().view(np.uint32), sparse.counts_host)


def test_assemble_follows_the_counted_sequence(cuda, tmp_path):
    """assemble_no_ends (reference assemble.py:21-184) with a BMM whose prior is tiny on a table counted from one
    random sequence: every k-mer has a single successor, so generation from a seed reproduces the source sequence
    forwards and -- on the reverse complement -- backwards, in MAP and in sampled mode; the reverse-strand counts
    come either from the table (summarised with reverse=True) or from the counter's reverse-complement lookups."""
    from bear_b200 import assemble, dataloader, summarize
    rng = np.random.default_rng(3)
    lag = 12
    src = ''.join(rng.choice(list('ACGT'), size=600))
    fa = tmp_path / 'seeds.fa'
    fa.write_text('>seed0\n' + src[300:312] + '\n' + src[312:320] + '\n>seed1\n' + src[100:130] + '\n')
    want = [src[270:360], src[90:135]]
    both, _ = summarize.count_kmers([src], [0], lag, reverse=True)
    fwd, _ = summarize.count_kmers([src], [0], lag, reverse=False)
    for table, reverse, get_map in ((both, False, True), (both, False, False), (fwd, True, False)):
        data = dataloader.KmerDataset(table, 1000)
        out = tmp_path / ('out_%d_%d' % (reverse, get_map))
        gen, ent = assemble.assemble_no_ends(str(fa), [[30, 40], [10, 5]], 3, None, van=1e-7, lag=lag, alphabet_name='dna',
                                             data=data, reverse=reverse, get_map=get_map, seed=5, batch_size=4,
                                             save_folder=str(out))
        assert gen.shape == (2, 3)
        for row, w in zip(gen, want):
            assert all(s == w for s in row), (row, w)
        assert all(np.allclose(e, 0.0) for e in ent) and len(ent[0]) == 90 and len(ent[1]) == 45
        assert (out / 'seqs.fa').read_text().count('>') == 6
    # a flat model (large prior, no counts to speak of) generates diverse sequences: per-site entropy near log 4
    data = dataloader.KmerDataset(fwd, 1000)
    gen, ent = assemble.assemble_no_ends([src[:lag]], [[0, 30]], 200, None, van=1e6, lag=lag, alphabet_name='dna',
                                         data=data, reverse=True, seed=6)
    assert len(set(gen[0])) > 150 and np.all(ent[0][lag:] > 1.2) and np.allclose(ent[0][:lag], 0.0)


def test_lookup_counts_repeated_keys_and_cached_index(cuda):
    """A k-mer held by several rows (tables concatenated from several files) gets the sum of its rows, absent
    k-mers get zeros, and the sorted index is built once per resident table (no sort per query)."""
    from bear_b200 import dataloader, get_var_probs
    rng = np.random.default_rng(5)
    kmers = np.array(['ACGT', 'CCGT', 'ACGT', '[[AC', 'TTTT', 'ACGT', 'CCGT'])
    counts = rng.integers(0, 50, size=(len(kmers), 2, 5))
    table = dataloader.KmerTable.from_arrays(kmers, counts, 'dna')
    data = dataloader.KmerDataset(table, 4)
    query = np.array(['ACGT', 'GGGG', 'CCGT', '[[AC', 'TTTT'])
    got, found = get_var_probs.lookup_counts(data, query, 'dna')
    want = np.stack([counts[kmers == q].sum(axis=0) if (kmers == q).any() else np.zeros((2, 5)) for q in query])
    assert found.cpu().tolist() == [True, False, True, True, True]
    assert np.array_equal(got.cpu().numpy(), want.astype(np.float64))
    index = table.sorted_index()
    get_var_probs.lookup_counts(data, query[:2], 'dna')
    assert table.sorted_index()[0] is index[0]              # the same tensors: nothing was re-sorted
