"""GPU parity tests: the CUDA path (through the C-ABI) against the CPU oracle on identical inputs.

Tolerances (float64): log-likelihoods 1e-10 relative (north-star bar: 1e-6); gradients 1e-8 relative
to the largest gradient component of the tensor; decode / integer aggregates bit-exact.
"""
import numpy as np
import pytest
import torch

from conftest import YSD1, SPARSE

pytestmark = pytest.mark.gpu

LL_RTOL = 1e-10
GRAD_RTOL = 1e-8


def _oracle():
    from oracle import bear_oracle as O
    return O


def rel_err(got, want):
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    return np.max(np.abs(got - want)) / max(np.max(np.abs(want)), 1e-300)


def synth_table(K, lag, G, seed, dense=False, start_frac=0.05):
    """Seeded synthetic table (host): random k-mers, a slice with start-padded prefixes, sparse or
    dense counts including all-zero rows."""
    rng = np.random.default_rng(seed)
    codes = rng.integers(0, 4 ** lag, size=K, dtype=np.uint64)
    nstart = np.where(rng.random(K) < start_frac, rng.integers(1, lag + 1, size=K), 0).astype(np.uint64)
    for i in np.flatnonzero(nstart):
        codes[i] &= np.uint64((1 << (2 * (lag - int(nstart[i])))) - 1)
    codes |= nstart << np.uint64(58)
    if dense:
        tot = np.round(np.exp(rng.normal(np.log(300), 1.5, size=(K, G)))).astype(np.int64)
    else:
        tot = rng.poisson(2.0, size=(K, G)) * (rng.random((K, G)) < 0.8)
    p = rng.dirichlet(0.3 * np.ones(5), size=(K, G))
    counts = np.stack([[rng.multinomial(tot[i, g], p[i, g]) for g in range(G)] for i in range(K)])
    return codes, counts.astype(np.int64)


def make_dataset(codes, counts, lag, batch):
    from bear_b200 import dataloader as dl
    table = dl.KmerTable.from_arrays((codes, lag), counts, 'dna')
    return dl.KmerDataset(table, batch)


# ------------------------------------------------------------------------------------------------
def test_dataloader_golden_first_batch(cuda):
    """reference tests/test_dataloader.py:20-32"""
    from bear_b200 import dataloader
    data = dataloader.dataloader(YSD1, 'dna', 3, 3)
    kmers, counts = next(iter(data))
    assert np.all(kmers.numpy() == np.array([b'TAATC', b'CGGTC', b'ACGCT']))
    counts_real = [[[14837, 15127, 22260, 16279, 446], [5029, 5095, 7408, 5487, 134], [16, 16, 23, 17, 0]],
                   [[61890, 729, 39733, 35956, 1017], [20524, 239, 13199, 12046, 309], [69, 0, 45, 39, 0]],
                   [[13965, 23135, 73870, 37045, 1035], [4705, 7591, 24532, 12305, 385], [14, 25, 81, 39, 0]]]
    assert counts.dtype == torch.float64
    assert np.all(counts.cpu().numpy() == np.array(counts_real))
    assert len(list(data)) == 1365 / 3


def test_one_hot_bit_exact(cuda):
    from bear_b200 import core, dataloader
    O = _oracle()
    data = dataloader.dataloader(YSD1, 'dna', 2000, 3)
    kmers, _ = next(iter(data))
    got = core.tf_one_hot(kmers, 'dna').cpu()
    want = O.one_hot(kmers.numpy())
    assert torch.equal(got, want)
    got2 = core.tf_one_hot(['[[ACG', 'TTTTT'], 'dna').cpu()
    assert torch.equal(got2, O.one_hot(['[[ACG', 'TTTTT']))
    prot = ['ARND[', '[[CEQ', 'WYVXA']
    assert torch.equal(core.tf_one_hot(prot, 'prot').cpu(), O.one_hot(prot, 'prot'))


def test_bmm_likelihood_known_answer(cuda):
    """reference tests/test_dataloader.py:34-49 + BASELINE.md known answers"""
    from bear_b200 import dataloader
    O = _oracle()
    data = dataloader.dataloader(YSD1, 'dna', 2000, 3)
    alpha = np.array([0.1, 1., 10.])
    got = dataloader.bmm_likelihood(data.map(lambda kmers, counts: counts), alpha).numpy()
    want = np.array([[-1.5271257134588018e8, -1.5270905139595887e8, -1.5274538628200924e8],
                     [-5.0819958884193577e7, -5.0816105408838786e7, -5.0848897074304186e7],
                     [-1.6336475286893066e5, -1.6157309667778228e5, -1.7003849692050868e5]])
    assert np.allclose(got, want, rtol=1e-11, atol=0)
    _, counts = O.read_tsv(YSD1, 3)
    assert rel_err(got, O.bmm_likelihood(counts, alpha).numpy()) < LL_RTOL
    # minibatched accumulation gives the same table
    got2 = dataloader.bmm_likelihood(dataloader.dataloader(YSD1, 'dna', 100, 3), alpha).numpy()
    assert np.allclose(got2, want, rtol=1e-11, atol=0)


def test_bmm_and_evaluation_count_regimes(cuda):
    """Counts on both sides of the per-CTA tables (64), rows mixing table letters with large letters, priors below
    and above the range of the constant-a Stirling form (5 alpha < 64), and huge counts: BMM table and the
    no-conditioning evaluation against the oracle."""
    from bear_b200 import bear_net, dataloader as dl
    O = _oracle()
    rng = np.random.default_rng(21)
    K, lag = 4096, 6
    codes, _ = synth_table(K, lag, 1, seed=4)
    counts = np.zeros((K, 2, 5), dtype=np.int64)
    counts[:, 0] = rng.integers(0, 3, size=(K, 5)) * rng.choice([1, 31, 32, 63, 64, 65, 1000, 250000, 4000000000 // 3], size=(K, 5))
    counts[:, 1] = rng.poisson(1.0, size=(K, 5)) * rng.choice([1, 70], size=(K, 5))
    counts[::7] = 0
    data = make_dataset(codes, counts, lag, 1000)
    alpha = np.array([0.05, 1.0, 12.0, 13.5, 300.0])
    got = dl.bmm_likelihood(data, alpha).numpy()
    want = O.bmm_likelihood(counts.astype(np.float64), alpha).numpy()
    assert rel_err(got, want) <= LL_RTOL
    for van in ([0.5, 2.0, 12.7], [40.0]):
        for h in (0.05, 1.0, 100.0):
            out = bear_net.evaluation(data, -1, 0, 'dna', h, None, van, seed=-1)
            ref = _oracle_eval(codes, counts, lag, 0, -1, [h], van, None)
            assert rel_err(out[0].numpy(), ref[0].numpy()) <= LL_RTOL and rel_err(out[2].numpy(), ref[2].numpy()) <= LL_RTOL


def test_core_distributions(cuda):
    """reference tests/test_core.py:7-26,42-60 (broadcast conc [5, A+1] against counts [3, 5, A+1])"""
    from scipy.special import loggamma
    from bear_b200 import core
    rng = np.random.default_rng(0)
    shape, A = np.array([3, 5]), 4
    trans = rng.poisson(size=np.r_[shape, A + 1]).astype(float)
    total = trans.sum(-1)
    conc = rng.exponential(size=np.r_[shape[1], A + 1])
    sum_conc = conc.sum(-1)
    dist = core.tfpDirichletMultinomialPerm(total, conc)
    assert np.all(dist._sample_n(7).cpu().numpy() == np.zeros(np.r_[7, shape, A + 1]))
    assert np.all(dist.ml_output(seed=-1).cpu().numpy() == np.tile(np.argmax(conc, -1)[None], [3, 1]))
    assert np.all(dist.ml_output(seed=5).cpu().numpy() == np.tile(np.argmax(conc, -1)[None], [3, 1]))
    want = (np.sum(loggamma(conc + trans) - loggamma(conc), -1) - (loggamma(sum_conc + total) - loggamma(sum_conc)))
    assert np.allclose(dist.counts_log_prob(trans).cpu().numpy(), want, rtol=1e-12, atol=1e-13)
    probs = conc / sum_conc[:, None]
    mn = core.tfpMultinomialPerm(total, probs)
    assert np.all(mn.ml_output(seed=-1).cpu().numpy() == np.tile(np.argmax(probs, -1)[None], [3, 1]))
    assert np.allclose(mn.counts_log_prob(trans).cpu().numpy(), np.sum(np.log(probs) * trans, -1), rtol=1e-12)
    # real-valued (non-integer) counts go through the lgamma path
    frac = trans + 0.25
    want = (np.sum(loggamma(conc + frac) - loggamma(conc), -1) - (loggamma(sum_conc + frac.sum(-1)) - loggamma(sum_conc)))
    assert np.allclose(core.tfpDirichletMultinomialPerm(frac.sum(-1), conc).counts_log_prob(frac).cpu().numpy(), want, rtol=1e-11)


def test_tie_breaking_is_uniform(cuda):
    """reference tests/test_core.py:29-39,63-73: ties resolved uniformly at random"""
    from scipy import stats as st
    from bear_b200 import core
    n = 4000
    conc = np.tile(np.array([1, 0.5, 1.]), (n, 1))
    for dist in (core.tfpDirichletMultinomialPerm(np.ones(n), conc), core.tfpMultinomialPerm(np.ones(n), conc / 2.5)):
        out = dist.ml_output(seed=123).cpu().numpy()
        assert set(np.unique(out)) <= {0., 2.}
        assert np.abs(np.sum(out - 1) / np.sqrt(n)) < st.norm.ppf(0.9995)


def test_lgamma_difference_accuracy(cuda):
    """The integer-offset lgamma/digamma differences against 50-digit mpmath on a grid that crosses
    every branch (zero, rising factorial, shift + Stirling, direct Stirling)."""
    import mpmath as mp
    from bear_b200 import core
    mp.mp.dps = 50
    a_vals = [1e-7, 1.3e-3, 0.2, 0.99999, 1.0, 3.7, 9.99, 10.0, 57.3, 2.5e5 + 0.1, 3.9e9]
    c_vals = [0, 1, 2, 9, 10, 11, 12, 37, 1000, 254715, 4000000000]
    conc = np.array([[a, 1.0] for a in a_vals for _ in c_vals])
    val = np.array([[float(c), 0.0] for _ in a_vals for c in c_vals])
    conc_t = torch.tensor(conc, device=cuda, requires_grad=True)
    out = core.tfpDirichletMultinomialPerm(val.sum(-1), conc_t).counts_log_prob(val)
    out.sum().backward()
    got, grad = out.detach().cpu().numpy(), conc_t.grad.cpu().numpy()
    for i, (cc, vv) in enumerate(zip(conc, val)):
        a, b, c = mp.mpf(cc[0]), mp.mpf(cc[1]), mp.mpf(vv[0])
        want = (mp.loggamma(a + c) - mp.loggamma(a)) - (mp.loggamma(a + b + c) - mp.loggamma(a + b))
        dwant = (mp.digamma(a + c) - mp.digamma(a)) - (mp.digamma(a + b + c) - mp.digamma(a + b))
        scale = max(abs(mp.loggamma(a + c) - mp.loggamma(a)), abs(mp.loggamma(a + b + c) - mp.loggamma(a + b)), mp.mpf(1e-300))
        assert abs(got[i] - want) <= 2e-14 * scale + 1e-300, (cc, vv, got[i], want)
        dscale = max(abs(mp.digamma(a + c) - mp.digamma(a)), abs(mp.digamma(a + b + c) - mp.digamma(a + b)), mp.mpf(1e-300))
        assert abs(grad[i, 0] - dwant) <= 2e-14 * dscale + 1e-300, (cc, vv, grad[i, 0], dwant)


# ------------------------------------------------------------------------------------------------
def _check_train_step(cuda, codes, counts, lag, col, train_ar, h_signed, seed, num_kmers=None):
    from bear_b200 import _lib, dataloader as dl
    from bear_b200._lib import lib, check, ptr
    O = _oracle()
    K = len(codes)
    num_kmers = K if num_kmers is None else num_kmers
    table = dl.KmerTable.from_arrays((codes, lag), counts, 'dna')
    k, c = table.device_tensors()
    gen = torch.Generator().manual_seed(seed)
    mat = O.init_linear(lag, 4, gen)[0] * 8.0
    hs = torch.tensor(float(h_signed), dtype=torch.float64)
    flat = torch.zeros(2 + mat.numel(), dtype=torch.float64, device=cuda)
    ll = torch.empty(K, dtype=torch.float64, device=cuda)
    ws = torch.empty(lib.bear_workspace_doubles(K, lag, mat.numel()), dtype=torch.float64, device=cuda)
    scale = num_kmers / K
    mat_d, hs_d = mat.to(cuda), hs.to(cuda)
    check(lib.bear_linear_train_step(ptr(k), table.col_ptr(col), table.stride, 0, K, lag, ptr(mat_d),
                                     ptr(hs_d), scale, int(train_ar), ptr(flat), ptr(ll), ptr(ws), _lib.stream()))
    torch.cuda.synchronize()
    kmers = dl.decode_kmers(codes, lag, 'dna')
    oh = O.one_hot(kmers)
    loss, ll_want, grads = O.train_step_grads(oh, torch.tensor(counts[:, col], dtype=torch.float64), hs, [mat],
                                              'linear', num_kmers, train_ar)
    flat = flat.cpu()
    assert abs(float(flat[0]) - float(loss)) <= LL_RTOL * abs(float(loss)), (float(flat[0]), float(loss))
    scale_ll = max(float(ll_want.abs().max()), 1e-300)
    assert float((ll.cpu() - ll_want).abs().max()) <= LL_RTOL * scale_ll
    assert abs(float(flat[1]) - float(grads[0])) <= GRAD_RTOL * max(abs(float(grads[0])), 1e-12 * abs(float(loss)))
    assert rel_err(flat[2:].numpy(), grads[1].reshape(-1).numpy()) <= GRAD_RTOL


@pytest.mark.parametrize('train_ar', [False, True])
@pytest.mark.parametrize('h_signed', [0.0, -3.1, 4.0])
def test_linear_train_step_ysd1(cuda, train_ar, h_signed):
    """C1: bundled lag-5 table (dense counts up to 2.5e5, 25 % start-padded k-mers), every column."""
    from bear_b200 import dataloader as dl
    t = dl.KmerTable.from_file(YSD1, 'dna', 3)
    codes = t.kmers_host[:t.num_rows].copy()
    counts = np.transpose(t.counts_host[:, :, :t.num_rows], (2, 0, 1)).astype(np.int64)
    for col in range(3):
        _check_train_step(cuda, codes, counts, 5, col, train_ar, h_signed, seed=col)


@pytest.mark.parametrize('lag,dense', [(1, False), (3, True), (4, False), (10, False), (13, True), (20, False), (29, False)])
def test_linear_train_step_synthetic(cuda, lag, dense):
    codes, counts = synth_table(3000, lag, 2, seed=lag, dense=dense)
    for train_ar in (False, True):
        _check_train_step(cuda, codes, counts, lag, 1, train_ar, -0.7, seed=lag, num_kmers=123456)


@pytest.mark.parametrize('lag', [2, 5, 13, 20])
def test_linear_train_step_mostly_start_padded(cuda, lag):
    """Half of the k-mers carry a start-symbol prefix of every possible length: they take the same chunk-table
    path as plain k-mers (the tables hold the start patterns), including the gradient rows of the start symbol."""
    codes, counts = synth_table(4000, lag, 1, seed=100 + lag, start_frac=0.5)
    _check_train_step(cuda, codes, counts, lag, 0, False, 0.4, seed=lag)
    _check_train_step(cuda, codes, counts, lag, 0, True, 0.4, seed=lag)


def test_linear_train_step_many_tiles_per_cta(cuda):
    """More rows than one pass of the persistent grid (148 CTAs x 22 tiles x 32 rows): every CTA iterates, the
    double-buffered stage is reused and the last iteration is ragged."""
    codes, counts = synth_table(148 * 22 * 32 * 2 + 777, 20, 1, seed=5)
    _check_train_step(cuda, codes, counts, 20, 0, False, -0.2, seed=1)


def test_linear_train_step_few_distinct_kmers(cuda):
    """Tiles in which many rows share chunk keys (a handful of distinct k-mers, repeated and in runs): the ranked
    read-modify-write rounds and the butterfly path for keys with more than five rows in a tile."""
    rng = np.random.default_rng(8)
    base, counts = synth_table(6, 9, 1, seed=2, start_frac=0.3)
    pick = np.concatenate([rng.integers(0, 6, size=2000), np.repeat(np.arange(6), 40), np.zeros(100, dtype=np.int64)])
    codes = base[pick]
    cnt = rng.poisson(2.0, size=(len(codes), 1, 5)).astype(np.int64)
    _check_train_step(cuda, codes, cnt, 9, 0, False, 0.1, seed=4)
    _check_train_step(cuda, codes, cnt, 9, 0, True, 0.1, seed=4)


@pytest.mark.parametrize('lag', [13, 20])
def test_linear_train_step_runs_of_equal_keys(cuda, lag):
    """Rows in k-mer order share their leading chunk keys in runs: the consumer's segmented-scan path (one
    read-modify-write per run).  Stray rows (out of order, start-padded) break the monotone order inside a tile, so
    run ends with possibly equal keys take turns; zero-count rows sit inside runs; the trailing chunks have 32 keys
    per tile and stay on the ranked path."""
    rng = np.random.default_rng(40 + lag)
    K = 5000
    _, counts = synth_table(K, lag, 1, seed=lag)
    prefix = np.uint64(int(rng.integers(0, 4 ** (lag - 11))) << 22)
    codes = np.sort(rng.choice(1 << 22, size=K, replace=False)).astype(np.uint64) | prefix
    stray = rng.choice(K, size=K // 40, replace=False)
    codes[stray] = rng.integers(0, 4 ** lag, size=len(stray), dtype=np.uint64)
    padded = rng.choice(K, size=K // 100, replace=False)
    for i in padded:
        ns = int(rng.integers(1, lag + 1))
        codes[i] = (codes[i] & np.uint64((1 << (2 * (lag - ns))) - 1)) | (np.uint64(ns) << np.uint64(58))
    for train_ar in (False, True):
        _check_train_step(cuda, codes, counts, lag, 0, train_ar, 0.3, seed=lag)
    # the same rows in sorted order without strays: monotone tiles only
    plain = np.sort(rng.choice(1 << 22, size=K, replace=False)).astype(np.uint64) | prefix
    _check_train_step(cuda, plain, counts, lag, 0, False, -0.4, seed=lag + 1)


def test_linear_head_exact_fallback_on_extreme_logits(cuda):
    """Logit spreads of several hundred overflow the product of table ratios; the kernels then evaluate the
    max-subtracted softmax of that row directly.  Train step and evaluation stay finite and match the oracle."""
    from bear_b200 import _lib, dataloader as dl
    from bear_b200._lib import lib, check, ptr
    O = _oracle()
    lag, K = 20, 600
    codes, counts = synth_table(K, lag, 1, seed=12, start_frac=0.1)
    gen = torch.Generator().manual_seed(3)
    mat = torch.randn(lag, 5, 5, dtype=torch.float64, generator=gen) * 60.0
    table = dl.KmerTable.from_arrays((codes, lag), counts, 'dna')
    k, c = table.device_tensors()
    hs = torch.tensor(0.2, dtype=torch.float64)
    flat = torch.zeros(2 + mat.numel(), dtype=torch.float64, device=cuda)
    ll = torch.empty(K, dtype=torch.float64, device=cuda)
    ws = torch.empty(lib.bear_workspace_doubles(K, lag, mat.numel()), dtype=torch.float64, device=cuda)
    mat_d, hs_d = mat.to(cuda), hs.to(cuda)
    check(lib.bear_linear_train_step(ptr(k), table.col_ptr(0), table.stride, 0, K, lag, ptr(mat_d), ptr(hs_d), 1.0, 0,
                                     ptr(flat), ptr(ll), ptr(ws), _lib.stream()))
    oh = O.one_hot(dl.decode_kmers(codes, lag, 'dna'))
    loss, ll_want, grads = O.train_step_grads(oh, torch.tensor(counts[:, 0], dtype=torch.float64), hs, [mat], 'linear', K, False)
    assert torch.isfinite(flat).all()
    assert abs(float(flat[0]) - float(loss)) <= 1e-9 * abs(float(loss))
    assert float((ll.cpu() - ll_want).abs().max()) <= 1e-9 * float(ll_want.abs().max())
    assert rel_err(flat[2:].cpu().numpy(), grads[1].reshape(-1).numpy()) <= 1e-7
    acc = torch.zeros(2 + 2 + 2 + 1, dtype=torch.float64, device=cuda)
    h = torch.tensor([1.5], dtype=torch.float64, device=cuda)
    van = torch.tensor([1.0], dtype=torch.float64, device=cuda)
    check(lib.bear_eval_step(ptr(k), table.col_ptr(0), None, table.stride, 0, K, lag, _lib.HEAD_LINEAR, ptr(mat_d), ptr(h), 1,
                             ptr(van), 1, -1, 0, ptr(acc), ptr(ws), _lib.stream()))
    want = _oracle_eval(codes, counts, lag, 0, -1, [1.5], [1.0], mat)
    assert abs(float(acc[0]) - float(want[0][0])) <= 1e-9 * abs(float(want[0][0]))
    assert abs(float(acc[1]) - float(want[1])) <= 1e-9 * abs(float(want[1]))


def test_linear_train_step_edge_rows(cuda):
    """empty batch is a no-op; ragged n (not a multiple of anything); all-zero rows contribute 0."""
    from bear_b200 import _lib, dataloader as dl
    from bear_b200._lib import lib, check, ptr
    codes, counts = synth_table(1001, 7, 1, seed=3)
    counts[::3] = 0
    _check_train_step(cuda, codes, counts, 7, 0, False, 0.3, seed=9)
    table = dl.KmerTable.from_arrays((codes, 7), counts, 'dna')
    k, c = table.device_tensors()
    flat = torch.zeros(2 + 7 * 25, dtype=torch.float64, device=cuda)
    ws = torch.empty(lib.bear_workspace_doubles(0, 7, 0), dtype=torch.float64, device=cuda)
    z = torch.zeros(7 * 25 + 1, dtype=torch.float64, device=cuda)
    check(lib.bear_linear_train_step(ptr(k), table.col_ptr(0), table.stride, 5, 0, 7, ptr(z[1:]), ptr(z[:1]), 1.0, 0,
                                     ptr(flat), None, ptr(ws), _lib.stream()))
    assert float(flat.abs().sum()) == 0.0
    # bad arguments are reported, not executed
    rc = lib.bear_linear_train_step(ptr(k), table.col_ptr(0), table.stride, 0, 10, 40, ptr(z[1:]), ptr(z[:1]), 1.0, 0,
                                    ptr(flat), None, ptr(ws), _lib.stream())
    assert rc == -1 and b'lag' in lib.bear_last_error()


# ------------------------------------------------------------------------------------------------
def _oracle_eval(codes, counts, lag, test_col, train_col, h, van, mat):
    from bear_b200 import dataloader as dl
    O = _oracle()
    oh = O.one_hot(dl.decode_kmers(codes, lag, 'dna'))
    f = O.ar_linear(oh, [mat]) if mat is not None else torch.zeros(len(codes), 5, dtype=torch.float64)
    test = counts[:, test_col].astype(np.float64)
    train = counts[:, train_col].astype(np.float64) if train_col >= 0 else None
    return O.evaluation([(oh, f, test, train)], torch.tensor(h, dtype=torch.float64), np.asarray(van, dtype=np.float64))


@pytest.mark.parametrize('train_col', [-1, 0])
def test_evaluation_ysd1(cuda, train_col):
    """bear_net.evaluation on C1 vs the oracle and the BASELINE.md known answers (tests/test_run.py:26-30)."""
    from bear_b200 import ar_funcs, bear_net, dataloader as dl
    O = _oracle()
    data = dl.dataloader(YSD1, 'dna', 500, 3)
    t = data.table
    codes = t.kmers_host[:t.num_rows]
    counts = np.transpose(t.counts_host[:, :, :t.num_rows], (2, 0, 1)).astype(np.int64)
    torch.manual_seed(1)
    ar_func, params = ar_funcs.make_ar_func_linear(5, 4)
    params[0].mul_(10.0)
    van = np.array([0.1, 1.0, 10.0])
    test_col = 0 if train_col < 0 else 1
    got = bear_net.evaluation(data, train_col, test_col, 'dna', 0.37, ar_func, van, seed=-1)
    want = _oracle_eval(codes, counts, 5, test_col, train_col, 0.37, van, params[0].cpu())
    for g, w in zip(got, want):
        assert rel_err(g.numpy(), w.numpy()) <= LL_RTOL
    if train_col < 0:
        assert np.allclose(got[2].numpy(), [-152712571.34208855, -152709051.39618367, -152745386.2824309], rtol=1e-11)
        assert np.allclose(got[5].numpy(), [3.7914698222, 3.7913533527, 3.7925557890], rtol=1e-9)
    else:
        assert np.allclose(got[2].numpy(), [-50794519.020665385, -50794605.339695275, -50795495.2242855], rtol=1e-11)
        assert np.allclose(got[5].numpy(), [3.790636628, 3.790645212, 3.790733706], rtol=1e-9)
        assert np.allclose(got[8].numpy(), 0.3676534498, rtol=1e-9)     # docs/usage.rst:255-265: 36.8 %


def test_evaluation_noise_only_moves_ties(cuda):
    """With the tie-breaking noise on, everything but the accuracies of tied rows is unchanged, and the
    no-conditioning BMM accuracy (all concentrations tied) is ~ the mean letter frequency."""
    from bear_b200 import ar_funcs, bear_net, dataloader as dl
    data = dl.dataloader(YSD1, 'dna', 1500, 3)
    torch.manual_seed(1)
    ar_func, params = ar_funcs.make_ar_func_linear(5, 4)
    a = bear_net.evaluation(data, -1, 0, 'dna', 1.0, ar_func, [1.0], seed=-1)
    b = bear_net.evaluation(data, -1, 0, 'dna', 1.0, ar_func, [1.0], seed=7)
    for i in (0, 1, 2, 3, 4, 5):
        assert np.array_equal(a[i].numpy(), b[i].numpy())
    assert 0.1 < float(b[8][0]) < 0.3          # random guess among 5 letters, stop is rare
    assert abs(float(a[6]) - float(b[6])) < 1e-3


def test_near_tie_noise_has_the_gaussian_law(cuda):
    """argmax(v + sigma N(0,1)) on rows whose two best entries differ by a fraction of sigma (core.py:134-136, AR model:
    sigma = eps): the better entry wins with probability Phi(gap / (sigma sqrt 2)); a third entry within the window is
    handled by the general path.  Entries further than 8 sigma away never win."""
    from scipy.stats import norm
    from bear_b200 import _lib
    from bear_b200._lib import lib, check, ptr
    n, eps = 400000, 1e-7
    dev = torch.device('cuda', 0)
    stride = n
    counts = torch.zeros((1, 5, stride), dtype=torch.int32, device=dev)
    counts[0, 0] = 1                                         # every row: one transition to letter 0
    ws = torch.empty(lib.bear_workspace_doubles(n, 1, 0), dtype=torch.float64, device=dev)
    h = torch.ones(1, dtype=torch.float64, device=dev)
    van = torch.ones(1, dtype=torch.float64, device=dev)
    for gap, third, want in ((0.5 * eps, 0.0, norm.cdf(0.5 / np.sqrt(2))), (1.5 * eps, 0.0, norm.cdf(1.5 / np.sqrt(2))),
                             (20 * eps, 0.0, 1.0), (0.5 * eps, 1.0, None)):
        f = torch.zeros((n, 5), dtype=torch.float64, device=dev)
        f[:, 0], f[:, 1] = 0.4, 0.4 - gap
        f[:, 2] = 0.4 - gap if third else 0.1                # third: a three-way near tie (general path)
        f[:, 3] = 1.0 - f[:, :3].sum(1)
        acc = torch.zeros(7, dtype=torch.float64, device=dev)
        check(lib.bear_eval_step(None, ptr(counts), None, stride, 0, n, 1, _lib.HEAD_EXPLICIT, ptr(f), ptr(h), 1, ptr(van), 1,
                                 11, 0, ptr(acc), ptr(ws), _lib.stream()))
        frac = float(acc[4]) / n                             # [ll_ear, ll_arm, ll_van, cor_ear, cor_arm, cor_van, total]
        if want is None:
            # P(x0 + d is the largest of three), x1 = x2 shifted down by d = 0.5 sigma: by simulation of the same law
            z = np.random.default_rng(0).normal(size=(2000000, 3))
            want = np.mean((z[:, 0] + 0.5 > z[:, 1]) & (z[:, 0] + 0.5 > z[:, 2]))
        assert abs(frac - want) < 4e-3, (gap, third, frac, want)


def test_h_scan_matches_evaluation(cuda):
    from bear_b200 import ar_funcs, bear_net, dataloader as dl
    data = dl.dataloader(YSD1, 'dna', 400, 3)
    torch.manual_seed(2)
    ar_func, _ = ar_funcs.make_ar_func_linear(5, 4)
    hs = np.exp(np.linspace(-6, 4, 11))        # 11 values: more than BEAR_MAX_MODELS per launch
    ll, perp, acc = bear_net.h_scan(data, 0, 1, 'dna', hs, ar_func, seed=-1)
    for i in (0, 5, 10):
        e = bear_net.evaluation(data, 0, 1, 'dna', hs[i], ar_func, [1.0], seed=-1)
        assert abs(float(ll[i]) - float(e[0])) <= 1e-12 * abs(float(e[0]))
        assert abs(float(perp[i]) - float(e[3])) <= 1e-12
        assert float(acc[i]) == float(e[6])


def test_evaluation_synthetic_sparse(cuda):
    codes, counts = synth_table(5000, 13, 3, seed=5)
    from bear_b200 import ar_funcs, bear_net
    data = make_dataset(codes, counts, 13, 1024)
    torch.manual_seed(3)
    ar_func, params = ar_funcs.make_ar_func_linear(13, 4)
    params[0].mul_(20.0)
    van = [0.5, 2.0]
    got = bear_net.evaluation(data, 2, 0, 'dna', 0.05, ar_func, van, seed=-1)
    want = _oracle_eval(codes, counts, 13, 0, 2, 0.05, van, params[0].cpu())
    for g, w in zip(got, want):
        assert rel_err(g.numpy(), w.numpy()) <= LL_RTOL


# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('train_ar', [False, True])
def test_training_trajectory_matches_oracle(cuda, train_ar):
    """A few Keras-Adam steps with gradient accumulation: parameters after training and the recorded
    losses follow the oracle's trajectory (bear_net.py:293-315)."""
    from bear_b200 import ar_funcs, bear_net, dataloader as dl
    O = _oracle()
    data = dl.dataloader(YSD1, 'dna', 300, 3)            # 5 batches, last one ragged (165 rows)
    K = data.table.num_rows
    torch.manual_seed(4)
    p0, h0, _ = bear_net._create_params(5, 4, ar_funcs.make_ar_func_linear, {})
    p0 = [p.clone() for p in p0]
    loss_save = []
    params, h_signed, ar_func = bear_net.train(data.repeat(3), K, 3, 0, 'dna', 5, ar_funcs.make_ar_func_linear, {},
                                               0.01, 'Adam', train_ar, acc_steps=2, params_restart=p0,
                                               loss_save=loss_save)
    kmers, counts = O.read_tsv(YSD1, 3)
    oh, c0 = O.one_hot(kmers), torch.tensor(counts[:, 0])
    batches = [(oh[i:i + 300], c0[i:i + 300]) for i in range(0, K, 300)] * 3
    want_loss = []
    wp, wh = O.train(batches, K, 'linear', [p0[1].cpu()], p0[0].cpu(), 0.01, train_ar, acc_steps=2, loss_save=want_loss)
    assert len(loss_save) == len(want_loss) == 7
    assert rel_err(loss_save, want_loss) <= 1e-9
    assert abs(float(h_signed) - float(wh)) <= 1e-9
    assert rel_err(params[1].cpu().numpy(), wp[0].numpy()) <= 1e-8
    assert params[0] is h_signed or float(params[0]) == float(h_signed)


def test_explicit_head_matches_fused_linear(cuda):
    """A plugin head (plain torch callable) through the explicit path reproduces the fused linear path."""
    from bear_b200 import ar_funcs, bear_net, dataloader as dl

    def make_ar_func_plugin(lag, alphabet_size, dtype=torch.float64):
        f, p = ar_funcs.make_ar_func_linear(lag, alphabet_size, dtype=dtype)
        return (lambda x: torch.softmax(torch.einsum('...jk,jkl->...l', x, p[0]), -1)), p

    data = dl.dataloader(YSD1, 'dna', 700, 3)
    K = data.table.num_rows
    torch.manual_seed(5)
    p0, _, _ = bear_net._create_params(5, 4, ar_funcs.make_ar_func_linear, {})
    p0 = [p.clone() for p in p0]
    la, lb = [], []
    pa, ha, fa = bear_net.train(data.repeat(2), K, 2, 0, 'dna', 5, ar_funcs.make_ar_func_linear, {}, 0.01, 'Adam',
                                False, params_restart=p0, loss_save=la)
    pb, hb, fb = bear_net.train(data.repeat(2), K, 2, 0, 'dna', 5, make_ar_func_plugin, {}, 0.01, 'Adam',
                                False, params_restart=p0, loss_save=lb)
    assert rel_err(la, lb) <= 1e-11
    assert rel_err(pa[1].cpu().numpy(), pb[1].cpu().numpy()) <= 1e-9
    ea = bear_net.evaluation(data, 0, 1, 'dna', torch.exp(ha), fa, [1.0], seed=-1)
    eb = bear_net.evaluation(data, 0, 1, 'dna', torch.exp(hb), fb, [1.0], seed=-1)
    for x, y in zip(ea, eb):
        assert rel_err(x.numpy(), y.numpy()) <= 1e-9


def test_cnn_head_train_step_matches_oracle(cuda):
    """CNN head (config bear_cnn_bear.cfg: filter_width 3) one accumulated step vs oracle autograd."""
    from bear_b200 import ar_funcs, bear_net, dataloader as dl
    O = _oracle()
    data = dl.dataloader(YSD1, 'dna', 1500, 3)
    K = data.table.num_rows
    torch.manual_seed(6)
    kw = {'filter_width': 3}
    p0, _, _ = bear_net._create_params(5, 4, ar_funcs.make_ar_func_cnn, kw)
    p0 = [p.clone() for p in p0]
    ls = []
    params, h_signed, ar_func = bear_net.train(data, K, 1, 0, 'dna', 5, ar_funcs.make_ar_func_cnn, kw, 0.01, 'SGD',
                                               False, params_restart=p0, loss_save=ls)
    kmers, counts = O.read_tsv(YSD1, 3)
    loss, _, grads = O.train_step_grads(O.one_hot(kmers), torch.tensor(counts[:, 0]), p0[0].cpu(),
                                        [p.cpu() for p in p0[1:]], 'cnn', K, False)
    assert abs(-ls[0] - float(loss)) <= 1e-10 * abs(float(loss))
    for new, old, g in zip(params, p0, grads):          # SGD: new = old - lr * grad
        got_g = (old.cpu() - new.cpu()) / 0.01
        assert rel_err(got_g.numpy(), g.numpy()) <= 1e-7


def test_cuda_graph_replay_matches_eager_loop(cuda, monkeypatch):
    """Many epochs over a small table are replayed from a CUDA graph; the trajectory equals the eager loop
    and the oracle's."""
    import time
    from bear_b200 import ar_funcs, bear_net, dataloader as dl
    O = _oracle()
    data = dl.dataloader(YSD1, 'dna', 700, 3)              # 2 batches per epoch
    K = data.table.num_rows
    torch.manual_seed(21)
    p0, _, _ = bear_net._create_params(5, 4, ar_funcs.make_ar_func_linear, {})
    p0 = [p.clone() for p in p0]
    runs = {}
    monkeypatch.setenv('BEAR_GRAPH_MIN_EPOCHS', '256')
    for mode in ('graph', 'eager'):
        if mode == 'eager':
            monkeypatch.setenv('BEAR_NO_GRAPH', '1')
        ls = []
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        params, h_signed, _ = bear_net.train(data.repeat(300), K, 300, 0, 'dna', 5, ar_funcs.make_ar_func_linear, {}, 0.01,
                                             'Adam', False, acc_steps=2, params_restart=p0, loss_save=ls)
        torch.cuda.synchronize()
        runs[mode] = (params[1].cpu(), float(h_signed), ls, time.perf_counter() - t0)
    g, e = runs['graph'], runs['eager']
    assert len(g[2]) == len(e[2]) == 300
    assert rel_err(g[2], e[2]) <= 1e-10 and rel_err(g[0].numpy(), e[0].numpy()) <= 1e-8 and abs(g[1] - e[1]) <= 1e-9
    kmers, counts = O.read_tsv(YSD1, 3)
    oh, c0 = O.one_hot(kmers), torch.tensor(counts[:, 0])
    batches = [(oh[i:i + 700], c0[i:i + 700]) for i in range(0, K, 700)] * 20
    wl = []
    O.train(batches, K, 'linear', [p0[1].cpu()], p0[0].cpu(), 0.01, False, acc_steps=2, loss_save=wl)
    assert rel_err(g[2][:20], wl) <= 1e-9
    print('300 epochs x 2 batches: graph %.3f s, eager %.3f s' % (g[3], e[3]))
