"""The BASELINE.json configurations as parity-test cases at sizes the CPU oracle finishes in seconds
(C1 is tests/test_gpu_parity.py + test_gpu_api.py on the bundled table; C5 is tests/test_gpu_properties.py
and the bench's own check)."""
import numpy as np
import pytest
import torch

from test_gpu_properties import synth

pytestmark = pytest.mark.gpu


def _host(table):
    k, c = table.device_tensors()
    K = table.num_rows
    kmers = [s.decode() for s in table.kmers_str()]
    counts = c[:, :, :K].permute(2, 0, 1).cpu().numpy().astype(np.float64)
    return kmers, counts


def rel(got, want):
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    return np.max(np.abs(got - want)) / max(np.max(np.abs(want)), 1e-300)


def test_c2_bear_ref_stop_head_eval_only(cuda):
    """C2: bear_ref empirical-transition BEAR (stop net, no AR training), all 4^lag k-mers enumerated, one data
    group + a reference column, evaluation-only log-likelihood / perplexity.  lag 6 against the oracle; the
    full lag-10 table (4^10 rows) through the closed-form tie of the BMM column."""
    from bear_b200 import ar_funcs, bear_ref, dataloader as dl
    from oracle import bear_oracle as O
    lag = 6
    K = 4 ** lag
    rng = np.random.default_rng(2)
    codes = np.arange(K, dtype=np.uint64)
    data_counts = np.round(np.exp(rng.normal(np.log(300), 1.5, size=(K, 1, 5)))).astype(np.int64)
    ref = (rng.random((K, 1, 5)) < 0.25).astype(np.int64) * rng.integers(1, 4, size=(K, 1, 5))
    ref[:, :, 4] = rng.integers(0, 2, size=(K, 1))          # stops in the reference are ignored (bear_ref.py:332-337)
    counts = np.concatenate([data_counts, ref], axis=1)
    table = dl.KmerTable.from_arrays((codes, lag), counts, 'dna')
    data = dl.KmerDataset(table, 1500)
    params, h_signed, ar_func = bear_ref._create_params(lag, 4, ar_funcs.make_ar_func_stop, {})
    van = np.array([0.1, 1.0, 10.0])
    got = bear_ref.evaluation(data, -1, 0, 1, 'dna', 0.0142, ar_func, van, seed=-1)
    kmers = [s.decode() for s in table.kmers_str()]
    oh = O.one_hot(kmers)
    f = O.ar_ref(oh, O.ref_counts_map(counts[:, 1].astype(float), 4), params[1].cpu(), params[2].cpu(),
                 lambda x: O.ar_stop(x, 4), 4)
    want = O.evaluation([(oh, f, counts[:, 0].astype(float), None)], torch.tensor(0.0142, dtype=torch.float64), van)
    for g, w in zip(got, want):
        assert rel(g.numpy(), w.numpy()) <= 1e-10
    # full C2 size, dense synthetic counts: the BMM column ties to the closed form
    big = synth(cuda, 4 ** 10, 10, 2, 1, start_permille=0)
    bdata = dl.KmerDataset(big, 1 << 18)
    p10, _, f10 = bear_ref._create_params(10, 4, ar_funcs.make_ar_func_stop, {})
    out = bear_ref.evaluation(bdata, -1, 0, 1, 'dna', 0.0142, f10, van, seed=1)
    assert np.allclose(out[2].numpy(), dl.bmm_likelihood(bdata, van + 1e-7)[0].numpy(), rtol=1e-12)
    assert np.all(np.isfinite([float(out[0]), float(out[1]), float(out[3]), float(out[4])]))


def test_c3_linear_lag13_eight_groups_fixed_steps(cuda):
    """C3: linear AR BEAR, lag 13, 8 groups, trained for a fixed number of steps on one column."""
    from bear_b200 import ar_funcs, bear_net, dataloader as dl
    from oracle import bear_oracle as O
    K, lag, G, col = 6000, 13, 8, 5
    table = synth(cuda, K, lag, G, 0)
    kmers, counts = _host(table)
    data = dl.KmerDataset(table, 2048)                       # 3 batches per epoch, last one ragged
    torch.manual_seed(8)
    p0, _, _ = bear_net._create_params(lag, 4, ar_funcs.make_ar_func_linear, {})
    p0 = [p.clone() for p in p0]
    ls = []
    params, h_signed, ar_func = bear_net.train(data.repeat(2), K, 2, col, 'dna', lag, ar_funcs.make_ar_func_linear, {},
                                               0.01, 'Adam', False, params_restart=p0, loss_save=ls)
    oh, c = O.one_hot(kmers), torch.tensor(counts[:, col])
    batches = [(oh[i:i + 2048], c[i:i + 2048]) for i in range(0, K, 2048)] * 2
    wl = []
    wp, wh = O.train(batches, K, 'linear', [p0[1].cpu()], p0[0].cpu(), 0.01, False, loss_save=wl)
    assert len(ls) == 6 and rel(ls, wl) <= 1e-9
    assert abs(float(h_signed) - float(wh)) <= 1e-9
    assert rel(params[1].cpu().numpy(), wp[0].numpy()) <= 1e-8
    # every group evaluates; the multi-dataset BMM table has one row per group
    bm = dl.bmm_likelihood(data, [1.0]).numpy()
    assert bm.shape == (8, 1) and rel(bm, O.bmm_likelihood(counts, np.array([1.0])).numpy()) <= 1e-10


def test_c4_cnn_lag13_four_groups_with_posterior_pass(cuda):
    """C4: CNN AR BEAR (filter_width 3, 30 filters, layer width 16), lag 13, 4 groups: train + eval + the
    get_pdf posterior pass (MC samples of the transition log-probabilities and their spread)."""
    from bear_b200 import ar_funcs, bear_net, dataloader as dl, get_var_probs
    from oracle import bear_oracle as O
    K, lag, G = 3000, 13, 4
    table = synth(cuda, K, lag, G, 1, start_permille=30)
    kmers, counts = _host(table)
    data = dl.KmerDataset(table, 1024)
    kw = {'filter_width': 3, 'num_filters': 30, 'kmer_layer1_width': 16}
    torch.manual_seed(9)
    p0, _, _ = bear_net._create_params(lag, 4, ar_funcs.make_ar_func_cnn, kw)
    p0 = [p.clone() for p in p0]
    assert sum(p.numel() for p in p0[1:]) == 6507
    ls = []
    params, h_signed, ar_func = bear_net.train(data, K, 1, 0, 'dna', lag, ar_funcs.make_ar_func_cnn, kw, 1e-3, 'Adam',
                                               False, params_restart=p0, loss_save=ls)
    oh, c = O.one_hot(kmers), torch.tensor(counts[:, 0])
    batches = [(oh[i:i + 1024], c[i:i + 1024]) for i in range(0, K, 1024)]
    wl = []
    wp, wh = O.train(batches, K, 'cnn', [p.cpu() for p in p0[1:]], p0[0].cpu(), 1e-3, False, loss_save=wl)
    assert rel(ls, wl) <= 1e-9
    for a, b in zip(params[1:], wp):
        assert rel(a.cpu().numpy(), b.numpy()) <= 1e-7
    h = float(torch.exp(h_signed))
    got = bear_net.evaluation(data, 0, 1, 'dna', h, ar_func, [0.1, 1.0], seed=-1)
    f = O.ar_cnn(oh, [p.cpu() for p in params[1:]])
    want = O.evaluation([(oh, f, counts[:, 1], counts[:, 0])], torch.tensor(h, dtype=torch.float64), np.array([0.1, 1.0]))
    for g, w in zip(got, want):
        assert rel(g.numpy(), w.numpy()) <= 1e-9
    # posterior pass: log p ~ log Dirichlet(conc); sample mean / variance of p against the closed forms
    sub = np.array(kmers[:200])
    mc = 4000
    lp = get_var_probs.get_pdf(sub, counts[:200], np.array([h]), ar_func, mc, [1.0], 0, 'dna', False, output='numpy', seed=4)
    assert lp.shape == (200, 5, 2, mc) and np.all(np.isfinite(lp))
    assert np.allclose(np.exp(lp).sum(1), 1.0, atol=1e-12)     # normalised over the letters
    ar_vals = f[:200].numpy()
    concs = O.get_pdf_concs(counts[:200, 0, :], ar_vals, np.array([h]), [1.0], False)     # [2, 200, 5]
    a0 = concs.sum(-1, keepdims=True)
    mean = np.transpose(concs / a0, (1, 2, 0))
    var = np.transpose(concs * (a0 - concs) / (a0 ** 2 * (a0 + 1)), (1, 2, 0))
    p = np.exp(lp)
    assert np.max(np.abs(p.mean(-1) - mean)) < 6 * np.sqrt(var.max() / mc) + 1e-3
    big = var > 1e-4
    assert np.allclose(p.var(-1)[big], var[big], rtol=0.25)


def test_protein_alphabet_dense_route(cuda):
    """Protein tables (A+1 = 21, 5-bit packing, unknown residue 'X' -> zero one-hot row) train and evaluate
    through the generic distribution kernels and match the oracle."""
    from bear_b200 import ar_funcs, bear_net, dataloader as dl
    from oracle import bear_oracle as O
    rng = np.random.default_rng(3)
    letters = O.ALPHABETS_IN['prot'][:-1]
    K, lag = 400, 3
    kmers = []
    for i in range(K):
        ns = int(rng.integers(0, lag + 1)) if rng.random() < 0.2 else 0
        body = list(rng.choice(letters, size=lag - ns))
        if body and rng.random() < 0.05:
            body[-1] = 'X'
        kmers.append('[' * ns + ''.join(body))
    counts = rng.poisson(1.5, size=(K, 2, 21)) * (rng.random((K, 2, 21)) < 0.4)
    table = dl.KmerTable.from_arrays(kmers, counts, 'prot')
    assert [k.decode() for k in table.kmers_str()] == kmers
    data = dl.KmerDataset(table, 150)
    torch.manual_seed(12)
    p0, _, _ = bear_net._create_params(lag, 20, ar_funcs.make_ar_func_linear, {})
    p0 = [p.clone() for p in p0]
    ls = []
    params, h_signed, ar_func = bear_net.train(data, K, 1, 0, 'prot', lag, ar_funcs.make_ar_func_linear, {}, 0.01, 'Adam',
                                               False, params_restart=p0, loss_save=ls)
    oh, c = O.one_hot(kmers, 'prot'), torch.tensor(counts[:, 0].astype(np.float64))
    batches = [(oh[i:i + 150], c[i:i + 150]) for i in range(0, K, 150)]
    wl = []
    wp, wh = O.train(batches, K, 'linear', [p0[1].cpu()], p0[0].cpu(), 0.01, False, loss_save=wl)
    assert rel(ls, wl) <= 1e-10
    assert rel(params[1].cpu().numpy(), wp[0].numpy()) <= 1e-8 and abs(float(h_signed) - float(wh)) <= 1e-9
    h = float(torch.exp(h_signed))
    got = bear_net.evaluation(data, 0, 1, 'prot', h, ar_func, [0.5, 2.0], seed=-1)
    f = O.ar_linear(oh, [params[1].cpu()])
    want = O.evaluation([(oh, f, counts[:, 1].astype(np.float64), counts[:, 0].astype(np.float64))],
                        torch.tensor(h, dtype=torch.float64), np.array([0.5, 2.0]))
    for g, w in zip(got, want):
        assert rel(g.numpy(), w.numpy()) <= 1e-10
    bm = dl.bmm_likelihood(data, [0.5, 2.0]).numpy()
    assert rel(bm, O.bmm_likelihood(counts.astype(np.float64), np.array([0.5, 2.0])).numpy()) <= 1e-10
    # h_scan (bear_net.py:465-531) on a protein table: every h agrees with evaluation() at that h
    hs = np.array([0.3 * h, h, 4.0 * h])
    ll, perp, accu = bear_net.h_scan(data, 0, 1, 'prot', hs, ar_func, seed=-1)
    for i, hk in enumerate(hs):
        one = bear_net.evaluation(data, 0, 1, 'prot', float(hk), ar_func, [1.0], seed=-1)
        assert rel(ll[i].numpy(), one[0].numpy()) <= 1e-12 and rel(perp[i].numpy(), one[3].numpy()) <= 1e-12
        assert rel(accu[i].numpy(), one[6].numpy()) <= 1e-12
