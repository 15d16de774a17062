"""Size-independent properties of the fused kernels on large synthetic tables (the BASELINE configs are
too big for the CPU oracle): additivity over row partitions, the closed-form BMM tie, zero-sum logit
gradients, exact integer aggregates, shard-regenerable synthetic data -- plus an oracle check on a
subsample of the same synthetic table."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def synth(cuda, K, lag, G, regime, seed=20, row_begin=0, start_permille=10):
    from bear_b200 import _lib, dataloader as dl
    from bear_b200._lib import lib, check, ptr
    stride = (K + 3) // 4 * 4
    kmers = torch.zeros(stride, dtype=torch.int64, device=cuda)
    counts = torch.zeros((G, 5, stride), dtype=torch.int32, device=cuda)
    check(lib.bear_synth_table(ptr(kmers), ptr(counts), stride, row_begin, K, lag, G, seed, regime, start_permille,
                               _lib.stream()))
    return dl.KmerTable.from_device(kmers, counts, K, lag, 'dna')


def train_step(table, col, r0, n, mat, hs, scale=1.0, train_ar=False):
    from bear_b200 import _lib
    from bear_b200._lib import lib, check, ptr
    k, c = table.device_tensors()
    flat = torch.zeros(2 + mat.numel(), dtype=torch.float64, device=k.device)
    ws = torch.empty(lib.bear_workspace_doubles(n, table.lag, mat.numel()), dtype=torch.float64, device=k.device)
    check(lib.bear_linear_train_step(ptr(k), table.col_ptr(col), table.stride, r0, n, table.lag, ptr(mat), ptr(hs), scale,
                                     int(train_ar), ptr(flat), None, ptr(ws), _lib.stream()))
    return flat


@pytest.mark.parametrize('lag,regime', [(20, 0), (13, 1)])
def test_train_step_is_additive_over_row_partitions(cuda, lag, regime):
    K = 1 << 21
    table = synth(cuda, K, lag, 2, regime)
    torch.manual_seed(0)
    mat = (torch.randn(lag, 5, 5, dtype=torch.float64, device=cuda) * 0.3).contiguous()
    hs = torch.tensor([-0.5], dtype=torch.float64, device=cuda)
    for train_ar in (False, True):
        whole = train_step(table, 1, 0, K, mat, hs, train_ar=train_ar)
        cut = 700001
        parts = train_step(table, 1, 0, cut, mat, hs, train_ar=train_ar) + train_step(table, 1, cut, K - cut, mat, hs, train_ar=train_ar)
        scale = whole.abs().max()
        assert float((whole - parts).abs().max()) <= 1e-11 * float(scale)
        # softmax backward: the five logit gradients of a row sum to zero, so does every [j, s, :] slice
        g = whole[2:].reshape(lag, 5, 5)
        assert float(g.sum(-1).abs().max()) <= 1e-9 * float(g.abs().max())
        if train_ar:
            assert float(whole[1]) == 0.0                   # no h gradient in AR mode (bear_net.py:193-196)
        # start symbols only occur in the padded 1 % slice: their weight rows still get gradient
        assert float(g[:, 4, :].abs().max()) > 0.0


def test_evaluation_ties_to_bmm_closed_form_at_scale(cuda):
    """tests/test_run.py:26-30 as a property: evaluation(ds_loc_train=-1).ll_van == bmm_likelihood(a + eps)."""
    from bear_b200 import ar_funcs, bear_net, dataloader as dl
    K = 1 << 22
    table = synth(cuda, K, 20, 2, 0)
    data = dl.KmerDataset(table, 1 << 20)
    torch.manual_seed(1)
    ar_func, _ = ar_funcs.make_ar_func_linear(20, 4)
    van = np.array([0.1, 1.0, 10.0])
    out = bear_net.evaluation(data, -1, 1, 'dna', 0.3, ar_func, van, seed=3)
    bmm = dl.bmm_likelihood(data, van + 1e-7)
    assert np.allclose(out[2].numpy(), bmm[1].numpy(), rtol=1e-12)
    k, c = table.device_tensors()
    total = int(c[1, :, :K].to(torch.int64).sum())
    assert float(out[0]) < 0 and float(out[1]) < 0
    # total_len is an exact integer aggregate: perplexity = exp(-ll / total_len)
    assert np.allclose(out[5].numpy(), np.exp(-out[2].numpy() / total), rtol=1e-14)
    # heldout: conditioning on column 0; accuracy of the BMM = mass of test counts on argmax of train counts
    held = bear_net.evaluation(data, 0, 1, 'dna', 0.3, ar_func, van, seed=-1)
    tr, te = c[0, :, :K].to(torch.int64), c[1, :, :K].to(torch.int64)
    best = tr.argmax(0)                                     # first maximum, like seed=-1
    want = int(te.gather(0, best[None]).sum()) / total
    assert np.allclose(held[8].numpy(), want, rtol=1e-14)


def test_zero_and_all_start_rows(cuda):
    """Rows without transitions contribute exactly 0; a table made only of start-padded k-mers takes the
    cooperative path for every row and still matches the additive / zero-sum properties."""
    K = 5000
    table = synth(cuda, K, 9, 1, 0, start_permille=1000)
    k, c = table.device_tensors()
    mat = (torch.randn(9, 5, 5, dtype=torch.float64, device=cuda) * 0.2).contiguous()
    hs = torch.zeros(1, dtype=torch.float64, device=cuda)
    a = train_step(table, 0, 0, K, mat, hs)
    b = train_step(table, 0, 0, 1234, mat, hs) + train_step(table, 0, 1234, K - 1234, mat, hs)
    assert float((a - b).abs().max()) <= 1e-11 * float(a.abs().max())
    c.zero_()
    z = train_step(table, 0, 0, K, mat, hs)
    assert float(z.abs().max()) == 0.0


def test_synthetic_table_is_shard_regenerable(cuda):
    whole = synth(cuda, 100000, 20, 2, 0)
    part = synth(cuda, 30000, 20, 2, 0, row_begin=50000)
    kw, cw = whole.device_tensors()
    kp, cp = part.device_tensors()
    assert torch.equal(kw[50000:80000], kp[:30000])
    assert torch.equal(cw[:, :, 50000:80000], cp[:, :, :30000])
    codes = kw[:100000] & ((1 << 58) - 1)
    plain = codes[(kw[:100000] >> 58) == 0]
    assert plain.unique().numel() == plain.numel()          # distinct k-mers


@pytest.mark.parametrize('regime', [0, 1])
def test_synthetic_subsample_matches_oracle(cuda, regime):
    """The first 4096 rows of the benchmark-style table against the CPU oracle (loss, gradients, evaluation)."""
    from bear_b200 import ar_funcs, bear_net, dataloader as dl
    from oracle import bear_oracle as O
    K, lag = 4096, 20
    table = synth(cuda, K, lag, 2, regime)
    k, c = table.device_tensors()
    kmers = [s.decode() for s in table.kmers_str()]
    counts = c[:, :, :K].permute(2, 0, 1).cpu().numpy().astype(np.float64)
    gen = torch.Generator().manual_seed(5)
    mat = O.init_linear(lag, 4, gen)[0] * 10
    hs = torch.tensor(0.7, dtype=torch.float64)
    oh = O.one_hot(kmers)
    for train_ar in (False, True):
        flat = train_step(table, 0, 0, K, mat.to(cuda), hs.reshape(1).to(cuda), scale=3.0, train_ar=train_ar).cpu()
        loss, _, grads = O.train_step_grads(oh, torch.tensor(counts[:, 0]), hs, [mat], 'linear', 3.0 * K, train_ar)
        assert abs(float(flat[0]) - float(loss)) <= 1e-10 * abs(float(loss))
        assert abs(float(flat[1]) - float(grads[0])) <= 1e-8 * max(abs(float(grads[0])), 1e-12 * abs(float(loss)))
        gw = grads[1].reshape(-1)
        assert float((flat[2:] - gw).abs().max()) <= 1e-8 * float(gw.abs().max())
    data = dl.KmerDataset(table, 1000)
    ar_func, params = ar_funcs.make_ar_func_linear(lag, 4)
    params[0].copy_(mat)
    got = bear_net.evaluation(data, 0, 1, 'dna', 2.0, ar_func, [0.1, 1.0], seed=-1)
    f = O.ar_linear(oh, [mat])
    want = O.evaluation([(oh, f, counts[:, 1], counts[:, 0])], torch.tensor(2.0, dtype=torch.float64), np.array([0.1, 1.0]))
    for g, w in zip(got, want):
        assert np.max(np.abs(g.numpy() - w.numpy())) <= 1e-10 * max(np.max(np.abs(w.numpy())), 1e-300)


def test_sorted_table_matches_oracle_and_is_additive(cuda):
    """Tables in KMC order (rows sorted by k-mer) make whole tiles share the leading chunk keys: the
    uniform-key path of the gradient scatter must agree with the oracle too."""
    from oracle import bear_oracle as O
    K, lag = 4096, 20
    table = synth(cuda, K, lag, 1, 2)
    k, c = table.device_tensors()
    codes = (k[:K] & ((1 << 58) - 1)).cpu().numpy()
    assert np.all(np.diff(codes[(k[:K] >> 58).cpu().numpy() == 0]) > 0)
    kmers = [s.decode() for s in table.kmers_str()]
    counts = c[:, :, :K].permute(2, 0, 1).cpu().numpy().astype(np.float64)
    gen = torch.Generator().manual_seed(6)
    mat = O.init_linear(lag, 4, gen)[0] * 10
    hs = torch.tensor(-0.2, dtype=torch.float64)
    flat = train_step(table, 0, 0, K, mat.to(cuda), hs.reshape(1).to(cuda)).cpu()
    loss, _, grads = O.train_step_grads(O.one_hot(kmers), torch.tensor(counts[:, 0]), hs, [mat], 'linear', K, False)
    assert abs(float(flat[0]) - float(loss)) <= 1e-10 * abs(float(loss))
    gw = grads[1].reshape(-1)
    assert float((flat[2:] - gw).abs().max()) <= 1e-8 * float(gw.abs().max())
    big = synth(cuda, 1 << 20, lag, 1, 2)
    a = train_step(big, 0, 0, 1 << 20, mat.to(cuda), hs.reshape(1).to(cuda))
    b = train_step(big, 0, 0, 300000, mat.to(cuda), hs.reshape(1).to(cuda)) + \
        train_step(big, 0, 300000, (1 << 20) - 300000, mat.to(cuda), hs.reshape(1).to(cuda))
    assert float((a - b).abs().max()) <= 1e-11 * float(a.abs().max())


@pytest.mark.parametrize('regime', [0, 1])
def test_bmm_and_evaluation_are_additive_over_ragged_row_ranges(cuda, regime):
    """The streaming kernels work on tiles aligned to absolute multiples of 32 / 128 rows fetched by the TMA engine; a
    batch may start and end anywhere.  Splitting a table at odd offsets (first tile partly dead, guarded tail tile,
    batches shorter than a tile) must give the same sums as one call, and the integer accuracy counts exactly."""
    from bear_b200 import _lib
    from bear_b200._lib import lib, check, ptr
    K, lag, G = 100_003, 11, 2
    table = synth(cuda, K, lag, G, regime)
    k, c = table.device_tensors()
    alpha = torch.tensor([0.3, 1.0, 7.0], dtype=torch.float64, device=cuda)
    ws = torch.empty(lib.bear_workspace_doubles(K, lag, 0), dtype=torch.float64, device=cuda)
    mat = (torch.randn(lag, 5, 5, dtype=torch.float64, device=cuda) * 0.3).contiguous()
    h = torch.tensor([0.7], dtype=torch.float64, device=cuda)

    def run(cuts):
        bmm = torch.zeros((G, 3), dtype=torch.float64, device=cuda)
        ev = torch.zeros(2 + 6 + 3, dtype=torch.float64, device=cuda)
        for lo, hi in zip(cuts[:-1], cuts[1:]):
            check(lib.bear_bmm_likelihood(ptr(c), table.stride, lo, hi - lo, G, 5, ptr(alpha), 3, ptr(bmm), ptr(ws), _lib.stream()))
            check(lib.bear_eval_step(ptr(k), table.col_ptr(0), table.col_ptr(1), table.stride, lo, hi - lo, lag, _lib.HEAD_LINEAR,
                                     ptr(mat), ptr(h), 1, ptr(alpha), 3, 99, lo, ptr(ev), ptr(ws), _lib.stream()))
        return bmm.cpu(), ev.cpu()
    whole_b, whole_e = run([0, K])
    parts_b, parts_e = run([0, 1, 17, 127, 129, 4097, 33333, 33334, 77777, K - 5, K])
    assert torch.allclose(whole_b, parts_b, rtol=1e-12, atol=0)
    assert torch.allclose(whole_e[:5], parts_e[:5], rtol=1e-12, atol=0)
    assert torch.equal(whole_e[5:], parts_e[5:])
