"""GPU parity of the fused CNN-head kernels (csrc/bear_cnn.cu, through the C-ABI) against the CPU oracle's
restatement of ar_funcs.make_ar_func_cnn (ar_funcs.py:49-99) + bear_net._train_step (bear_net.py:146-197).

Tolerances (float64): head outputs / log-likelihoods 1e-10 relative; gradients 1e-8 relative to the largest
component of each parameter tensor.
"""
import numpy as np
import pytest
import torch

from test_gpu_parity import rel_err, synth_table, make_dataset, _oracle

pytestmark = pytest.mark.gpu


def random_cnn_params(lag, W, F, H1, seed):
    """Oracle-initialised parameters with every intercept / scale perturbed, so no gradient is trivially 0."""
    O = _oracle()
    gen = torch.Generator().manual_seed(seed)
    ps = O.init_cnn(lag, 4, gen, filter_width=W, num_filters=F, kmer_layer1_width=H1)
    ps = [p + 0.3 * torch.randn(p.shape, dtype=torch.float64, generator=gen) if i in (1, 3, 5, 6, 7) else p
          for i, p in enumerate(ps)]
    ps[4] = ps[4] * 10.0          # a head that is not almost uniform
    return ps


def device_block(ps, dev):
    return torch.cat([p.reshape(-1) for p in ps]).to(dev).contiguous()


def split_block(flat, ps):
    out, o = [], 0
    for p in ps:
        out.append(flat[o:o + p.numel()].reshape(p.shape))
        o += p.numel()
    return out


DIMS = [            # lag, W, F, H1, rows
    (13, 3, 30, 16, 6000),      # BASELINE C4 / config_files/bear_cnn_bear.cfg; 32-row tiles, 16 warps; > 148 tiles
    (13, 8, 30, 16, 700),       # the reference's default filter width (ar_funcs.py:50)
    (5, 3, 30, 16, 333),        # bundled example lag
    (9, 2, 7, 5, 1000),         # odd sizes: F and H1 not multiples of the tensor-core tile
    (20, 8, 30, 16, 500),       # 16-row tiles, 8 warps (shared-memory bound)
    (6, 6, 32, 8, 200),         # a single conv position, all 32 lanes active
    (16, 8, 30, 16, 400),       # 32-row tiles with 8 warps (the 16 warp-private filter-gradient tables do not fit)
    (17, 3, 28, 16, 300),       # 16-row tiles with 16 warps (too many dW1 accumulator tiles for 8 warps)
]


@pytest.mark.parametrize('lag,W,F,H1,K', DIMS)
def test_cnn_forward_matches_oracle(cuda, lag, W, F, H1, K):
    from bear_b200 import _lib
    from bear_b200._lib import lib, check, ptr
    O = _oracle()
    assert lib.bear_cnn_supported(lag, W, F, H1) == 1
    codes, _ = synth_table(K, lag, 1, seed=lag * 100 + W, start_frac=0.2)
    ps = random_cnn_params(lag, W, F, H1, seed=3)
    assert lib.bear_cnn_num_params(lag, W, F, H1) == sum(p.numel() for p in ps)
    block = device_block(ps, cuda)
    k = torch.as_tensor(codes.view(np.int64), device=cuda)
    f = torch.empty((K, 5), dtype=torch.float64, device=cuda)
    check(lib.bear_cnn_head_forward(ptr(k), 0, K, lag, W, F, H1, ptr(block), ptr(f), _lib.stream()))
    from bear_b200 import dataloader as dl
    kmers = dl.KmerTable.from_arrays((codes, lag), np.zeros((K, 1, 5), dtype=np.int64), 'dna').kmers_str()
    want = O.ar_cnn(O.one_hot(kmers), ps)
    assert rel_err(f.cpu().numpy(), want.numpy()) <= 1e-12
    # a row offset addresses the same rows
    f2 = torch.empty((K - 37, 5), dtype=torch.float64, device=cuda)
    check(lib.bear_cnn_head_forward(ptr(k), 37, K - 37, lag, W, F, H1, ptr(block), ptr(f2), _lib.stream()))
    assert torch.equal(f2, f[37:])


@pytest.mark.parametrize('train_ar', [False, True])
@pytest.mark.parametrize('lag,W,F,H1,K', DIMS)
def test_cnn_train_step_matches_oracle_autograd(cuda, lag, W, F, H1, K, train_ar):
    from bear_b200 import _lib, dataloader as dl
    from bear_b200._lib import lib, check, ptr
    O = _oracle()
    codes, counts = synth_table(K, lag, 2, seed=lag * 7 + W, dense=(lag == 9), start_frac=0.1)
    table = dl.KmerTable.from_arrays((codes, lag), counts, 'dna')
    k, c = table.device_tensors()
    ps = random_cnn_params(lag, W, F, H1, seed=11)
    block = device_block(ps, cuda)
    h_signed = torch.tensor(0.3, dtype=torch.float64, device=cuda)
    nparams = block.numel()
    flat = torch.zeros(2 + nparams, dtype=torch.float64, device=cuda)
    ll = torch.empty(K, dtype=torch.float64, device=cuda)
    ws = torch.empty(lib.bear_workspace_doubles(K, lag, nparams), dtype=torch.float64, device=cuda)
    scale = 3.7
    check(lib.bear_cnn_train_step(ptr(k), table.col_ptr(1), table.stride, 0, K, lag, W, F, H1, ptr(block), ptr(h_signed),
                                  scale, int(train_ar), ptr(flat), ptr(ll), ptr(ws), _lib.stream()))
    oh = O.one_hot(table.kmers_str())
    c1 = torch.tensor(counts[:, 1]).to(torch.float64)
    loss, ll_want, grads = O.train_step_grads(oh, c1, h_signed.cpu(), ps, 'cnn', scale * K, train_ar)
    flat = flat.cpu()
    assert abs(float(flat[0]) - float(loss)) <= 1e-10 * abs(float(loss))
    assert rel_err(ll.cpu().numpy(), ll_want.numpy()) <= 1e-10
    gh = float(grads[0])
    assert abs(float(flat[1]) - gh) <= 1e-8 * max(abs(gh), 1e-300) or (train_ar and gh == 0.0 and float(flat[1]) == 0.0)
    for name, got, want in zip(['filters', 'int0', 'W1', 'int1', 'W2', 'int2', 'scale0', 'scale1'],
                               split_block(flat[2:], ps), grads[1:]):
        assert rel_err(got.numpy(), want.numpy()) <= 1e-8, name
    # adding into a non-zero buffer accumulates (acc_steps > 1, bear_net.py:193-196)
    flat2 = torch.ones(2 + nparams, dtype=torch.float64, device=cuda)
    check(lib.bear_cnn_train_step(ptr(k), table.col_ptr(1), table.stride, 0, K, lag, W, F, H1, ptr(block), ptr(h_signed),
                                  scale, int(train_ar), ptr(flat2), None, ptr(ws), _lib.stream()))
    assert rel_err((flat2.cpu() - 1.0).numpy(), flat.numpy()) <= 1e-9


def test_cnn_backward_matches_autograd(cuda):
    """bear_cnn_head_backward: parameter gradients of sum(gf * f) for an arbitrary upstream gf."""
    from bear_b200 import _lib
    from bear_b200._lib import lib, check, ptr
    from bear_b200 import dataloader as dl
    O = _oracle()
    lag, W, F, H1, K = 13, 3, 30, 16, 5000
    codes, _ = synth_table(K, lag, 1, seed=5, start_frac=0.1)
    table = dl.KmerTable.from_arrays((codes, lag), np.zeros((K, 1, 5), dtype=np.int64), 'dna')
    ps = random_cnn_params(lag, W, F, H1, seed=2)
    block = device_block(ps, cuda)
    gen = torch.Generator().manual_seed(9)
    gf = torch.randn(K, 5, dtype=torch.float64, generator=gen)
    k, _ = table.device_tensors()
    gp = torch.zeros_like(block)
    ws = torch.empty(lib.bear_workspace_doubles(K, lag, block.numel()), dtype=torch.float64, device=cuda)
    check(lib.bear_cnn_head_backward(ptr(k), 0, K, lag, W, F, H1, ptr(block), ptr(gf.to(cuda)), ptr(gp), ptr(ws),
                                     _lib.stream()))
    req = [p.clone().requires_grad_(True) for p in ps]
    f = O.ar_cnn(O.one_hot(table.kmers_str()), req)
    want = torch.autograd.grad((f * gf).sum(), req)
    for got, w in zip(split_block(gp.cpu(), ps), want):
        assert rel_err(got.numpy(), w.numpy()) <= 1e-8


def test_cnn_unsupported_shapes_are_refused(cuda):
    from bear_b200 import _lib
    from bear_b200._lib import lib, ptr
    assert lib.bear_cnn_supported(13, 3, 64, 16) == 0          # more filters than lanes
    assert lib.bear_cnn_supported(13, 3, 30, 32) == 0          # layer wider than a half-warp
    assert lib.bear_cnn_supported(29, 2, 32, 16) == 0          # tile does not fit shared memory
    assert lib.bear_cnn_supported(5, 6, 30, 16) == 0           # filter wider than the lag
    k = torch.zeros(8, dtype=torch.int64, device=cuda)
    f = torch.empty((8, 5), dtype=torch.float64, device=cuda)
    blk = torch.zeros(100000, dtype=torch.float64, device=cuda)
    assert lib.bear_cnn_head_forward(ptr(k), 0, 8, 13, 3, 64, 16, ptr(blk), ptr(f), _lib.stream()) == -5
    assert 'outside the fused kernel' in _lib.last_error()


def test_api_routes_cnn_through_fused_kernels_and_matches_torch_route(cuda, monkeypatch):
    """bear_net.train / evaluation and bear_ref.train with the CNN head: the fused kernels (default) and the
    torch-op explicit route (BEAR_CNN_TORCH=1, the plugin path) follow the same trajectory."""
    from bear_b200 import ar_funcs, bear_net, bear_ref, _engine as eng
    lag, K = 13, 4000
    codes, counts = synth_table(K, lag, 3, seed=77, start_frac=0.05)
    counts[:, 2, 4] = 0
    data = make_dataset(codes, counts, lag, 1500)
    kw = {'filter_width': 3}
    torch.manual_seed(1)
    p0, _, af0 = bear_net._create_params(lag, 4, ar_funcs.make_ar_func_cnn, kw)
    assert eng.fused_cnn_ok(af0, data.table)
    p0 = [p.clone() for p in p0]
    q0, _, _ = bear_ref._create_params(lag, 4, ar_funcs.make_ar_func_cnn, kw)
    q0 = [q.clone() for q in q0]
    res = {}
    for route in ('fused', 'torch'):
        if route == 'torch':
            monkeypatch.setenv('BEAR_CNN_TORCH', '1')
        ls, ls2 = [], []
        params, h_signed, ar_func = bear_net.train(data.repeat(2), K, 2, 0, 'dna', lag, ar_funcs.make_ar_func_cnn, kw,
                                                   1e-3, 'Adam', False, params_restart=p0, loss_save=ls)
        ev = bear_net.evaluation(data, 0, 1, 'dna', torch.exp(h_signed), ar_func, [0.5, 2.0], seed=-1)
        rp, rh, raf = bear_ref.train(data, K, 1, 0, 2, 'dna', lag, ar_funcs.make_ar_func_cnn, kw, 1e-3, 'Adam', False,
                                     params_restart=q0, loss_save=ls2)
        res[route] = (ls, [p.cpu() for p in params], [e.numpy() for e in ev], ls2, [p.cpu() for p in rp])
    a, b = res['fused'], res['torch']
    assert rel_err(a[0], b[0]) <= 1e-10 and rel_err(a[3], b[3]) <= 1e-10
    for x, y in zip(a[1] + a[4], b[1] + b[4]):
        assert rel_err(x.numpy(), y.numpy()) <= 1e-7
    for x, y in zip(a[2], b[2]):
        assert rel_err(x, y) <= 1e-9
