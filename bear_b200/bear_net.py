"""Train and evaluate BEAR / AR / BMM models on packed, device-resident k-mer count tables.

Mirrors the reference's ``bear_model/bear_net.py`` entry points (``train`` bear_net.py:200-321,
``evaluation`` :387-463, ``h_scan`` :465-531, ``change_scope_params`` :103-143 and the private
``_bear_kmer_counts`` / ``_ar_kmer_counts`` / ``_create_params`` / ``_train_step`` /
``_evaluation_step`` helpers).  ``data`` is a ``bear_b200.dataloader.KmerDataset`` instead of a
``tf.data`` object; everything else keeps the reference's argument names, order and meaning.

Hot path: with the built-in linear head on DNA/RNA each training batch is ONE fused kernel
(``bear_linear_train_step``: 2-bit decode -> table-gather head -> lgamma/digamma forward+backward ->
in-kernel reductions) that adds ``[loss, d h_signed, d mat]`` into a flat buffer; evaluation is
``bear_eval_step``.  The built-in CNN head is one fused kernel too (``bear_cnn_train_step``: conv as a
gather, dense layers on the FP64 tensor cores, loss and the whole backward pass on-chip).  Other heads
(user plugins, CNN shapes outside the fused kernel) are evaluated with torch ops on the device and enter
through ``bear_dm_train_step_explicit`` / ``BEAR_HEAD_EXPLICIT``.
"""
import numpy as np
import torch

from . import _engine as eng
from . import _lib, core
from ._lib import lib, check, ptr

epsilon = 1e-7    # tf.keras.backend.epsilon(), bear_net.py:4


def _bear_kmer_counts(kmer_seqs, kmer_total_counts, condition_trans_counts=None, h=None, ar_func=None):
    """Distribution of k-mer transition counts under a BEAR model (bear_net.py:7-45):
    concentrations = ar_func(kmer_seqs) / h + condition_trans_counts + epsilon."""
    dtype = kmer_seqs.dtype
    if condition_trans_counts is None:
        condition_trans_counts = torch.zeros((), dtype=dtype, device=kmer_seqs.device)
    if h is None or ar_func is None:
        h = torch.ones((), dtype=dtype, device=kmer_seqs.device)

        def ar_func(x):
            return torch.zeros((), dtype=dtype, device=kmer_seqs.device)
    concentrations = ar_func(kmer_seqs) / h + condition_trans_counts + epsilon
    return core.tfpDirichletMultinomialPerm(kmer_total_counts, concentrations, name='x')


def _ar_kmer_counts(kmer_seqs, kmer_total_counts, ar_func):
    """Distribution of k-mer transition counts under an AR model (bear_net.py:48-70):
    probs = ar_func(kmer_seqs) + epsilon (not renormalised)."""
    probs = ar_func(kmer_seqs) + epsilon
    return core.tfpMultinomialPerm(kmer_total_counts, probs, name='x')


def _create_params(lag, alphabet_size, make_ar_func, af_kwargs, dtype=torch.float64):
    """params = [h_signed (= 0)] + ar_func params (bear_net.py:73-100)."""
    ar_func, ar_func_params = make_ar_func(lag, alphabet_size, **af_kwargs, dtype=dtype)
    h_signed = torch.zeros((), dtype=dtype, device=_lib.device())
    params = [h_signed] + ar_func_params
    return params, h_signed, ar_func


def _param_value(p, dtype, device):
    if isinstance(p, torch.Tensor):
        return p.detach().to(device=device, dtype=dtype)
    if hasattr(p, 'numpy'):
        p = p.numpy()
    return torch.as_tensor(np.asarray(p), dtype=dtype, device=device)


def change_scope_params(lag, alphabet_size, make_ar_func, af_kwargs, params, dtype=torch.float64):
    """Unpack a parameter list ``[h_signed, *ar_params]`` (the model-file layout,
    models/train_bear_net.py:147-149) into fresh parameters and their ar_func (bear_net.py:103-143)."""
    new_params, h_signed, ar_func = _create_params(lag, alphabet_size, make_ar_func, af_kwargs, dtype=dtype)
    if len(params) != len(new_params):
        raise ValueError('expected %d parameters, got %d' % (len(new_params), len(params)))
    for dst, src in zip(new_params, params):
        dst.copy_(_param_value(src, dst.dtype, dst.device).reshape(dst.shape))
    return new_params, h_signed, ar_func


def _train_step(batch, num_kmers, h_signed, ar_func, params, acc_grads, train_ar):
    """Dense-tensor training step (bear_net.py:146-197) for callers that hold one-hot / count tensors:
    adds d loss / d params into ``acc_grads`` and returns the loss."""
    kmer_seqs, transition_counts = batch[0], batch[1]
    kmer_batch_size = kmer_seqs.shape[0]
    kmer_total_counts = transition_counts.sum(-1)
    was = [p.requires_grad for p in params]
    for p in params:
        p.requires_grad_(True)
    try:
        with torch.enable_grad():
            if train_ar:
                post = _ar_kmer_counts(kmer_seqs, kmer_total_counts, ar_func)
            else:
                post = _bear_kmer_counts(kmer_seqs, kmer_total_counts, h=torch.exp(h_signed), ar_func=ar_func)
            log_likelihood = post.counts_log_prob(transition_counts).sum()
            loss = -(num_kmers / kmer_batch_size) * log_likelihood
            grads = torch.autograd.grad(loss, params, allow_unused=True)
    finally:
        for p, w in zip(params, was):
            p.requires_grad_(w)
    for tv, g in zip(acc_grads, grads):
        if g is not None:
            tv.add_(g)
    return loss.detach()


def train(data, num_kmers, epochs, ds_loc, alphabet, lag, make_ar_func, af_kwargs,
          learning_rate, optimizer_name, train_ar, acc_steps=1,
          params_restart=None, writer=None, loss_save=None, dtype=torch.float64):
    """Train a BEAR or AR model (bear_net.py:200-321).

    data : KmerDataset, minibatched and already ``.repeat(epochs)``-ed (``epochs`` itself is unused,
        as in the reference).  num_kmers : total rows, normalises the loss estimate.  ds_loc : count
        column to train on.  make_ar_func / af_kwargs : the AR-head plugin.  optimizer_name : Keras
        optimizer name (Adam, SGD, RMSprop, Adagrad).  train_ar : AR (True) or BEAR (False) likelihood.
        acc_steps : batches to accumulate (sum) before each update.  params_restart : parameter list
        of a previous run.  writer : object with ``add_scalar`` (optional).  loss_save : list that
        receives -loss/acc_steps per update.
    Returns (params, h_signed, ar_func) with params = [h_signed, *ar_params] (torch CUDA tensors).
    """
    if dtype not in (torch.float64, 'float64'):
        raise ValueError('bear_b200 computes in float64 (the reference default and recommendation)')
    table = eng.check_dataset(data)
    alphabet_size = len(core.alphabets_tf[alphabet]) - 1
    if params_restart is None:
        params, h_signed, ar_func = _create_params(lag, alphabet_size, make_ar_func, af_kwargs)
    else:
        params, h_signed, ar_func = change_scope_params(lag, alphabet_size, make_ar_func, af_kwargs, params_restart)
    fp = eng.FlatParams(params)
    ws = eng.workspace(table, fp.total)
    k, c = table.device_tensors()

    graph_safe = False
    if table.A1 != 5:
        # protein tables: dense route through the generic distribution kernels (any alphabet size)
        def step_fn(r0, n, scale):
            for c0 in range(r0, r0 + n, eng.EXPLICIT_CHUNK):
                cn = min(eng.EXPLICIT_CHUNK, r0 + n - c0)
                batch = (eng.onehot_rows(table, c0, cn), eng.dense_counts(table, c0, cn)[:, ds_loc, :])
                grads = [fp.grad_view(i) for i in range(len(params))]
                fp.grad[0] += _train_step(batch, scale * cn, h_signed, ar_func, params, grads, train_ar)
    elif eng.fused_linear_ok(ar_func, table):
        mat = params[1]
        graph_safe = True            # the step is libbear_b200 launches only: capturable in a CUDA graph

        def step_fn(r0, n, scale):
            check(lib.bear_linear_train_step(ptr(k), table.col_ptr(ds_loc), table.stride, r0, n, table.lag,
                                             ptr(mat), ptr(h_signed), scale, int(train_ar), ptr(fp.grad), None,
                                             ptr(ws), _lib.stream()))
    elif eng.fused_cnn_ok(ar_func, table):
        W, F, H1 = eng.cnn_dims(ar_func)
        block = eng.cnn_param_block(params[1:])      # views of fp.flat: contiguous, in the reference order
        assert block.data_ptr() == params[1].data_ptr()
        graph_safe = True

        def step_fn(r0, n, scale):
            check(lib.bear_cnn_train_step(ptr(k), table.col_ptr(ds_loc), table.stride, r0, n, table.lag, W, F, H1,
                                          ptr(block), ptr(h_signed), scale, int(train_ar), ptr(fp.grad), None,
                                          ptr(ws), _lib.stream()))
    else:
        for p in params[1:]:
            p.requires_grad_(True)

        def step_fn(r0, n, scale):
            eng.explicit_train_step(table, ds_loc, r0, n, scale, train_ar, fp, ws,
                                    lambda c0, cn: eng.explicit_f(ar_func, table, c0, cn))

    eng.train_loop(data, num_kmers, ds_loc, train_ar, acc_steps, fp, optimizer_name, learning_rate, step_fn,
                   writer=writer, loss_save=loss_save, graph_safe=graph_safe and optimizer_name == 'Adam')
    for p in params:
        p.requires_grad_(False)
    return params, h_signed, ar_func


def _head_for_eval(ar_func, table):
    """(BEAR_HEAD id, head_ptr_fn) for bear_eval_step."""
    kind = eng.head_kind(ar_func)
    if kind == 'none':
        return _lib.HEAD_NONE, (lambda r0, n: (None, None))
    if kind == 'stop':
        return _lib.HEAD_STOP, (lambda r0, n: (None, None))
    if eng.fused_linear_ok(ar_func, table):
        mat = ar_func.params[0].detach().contiguous()
        return _lib.HEAD_LINEAR, (lambda r0, n: (mat, ptr(mat)))
    if eng.fused_cnn_ok(ar_func, table):
        block = eng.cnn_param_block(ar_func.params)
        return _lib.HEAD_EXPLICIT, eng.explicit_head_ptr_fn(lambda c0, cn: eng.cnn_forward(ar_func, table, c0, cn, block))
    return _lib.HEAD_EXPLICIT, eng.explicit_head_ptr_fn(lambda c0, cn: eng.explicit_f(ar_func, table, c0, cn))


def _evaluation_step(batch, h, ar_func, van_reg, alphabet_size, use_train, dtype=torch.float64, seed=None):
    """Dense-tensor evaluation step (bear_net.py:323-371) on (one-hot, test counts[, train counts])."""
    kmer_seqs, test = batch[0], batch[1]
    van_reg = torch.as_tensor(np.asarray(van_reg, dtype=np.float64)).to(test.device)
    h = h if isinstance(h, torch.Tensor) else torch.as_tensor(np.asarray(h, dtype=np.float64))
    h = h.to(test.device)
    if use_train:
        train_c = batch[2]
        van_condition = train_c[:, None, :] + van_reg[:, None]
    else:
        train_c = None
        van_condition = van_reg[:, None] * torch.ones((1, alphabet_size + 1), dtype=torch.float64, device=test.device)
    tot = test.sum(-1)
    post_ear = _bear_kmer_counts(kmer_seqs, tot, condition_trans_counts=train_c, h=h, ar_func=ar_func)
    post_arm = _ar_kmer_counts(kmer_seqs, tot, ar_func)
    post_van = _bear_kmer_counts(kmer_seqs, tot[:, None], condition_trans_counts=van_condition)
    ll_ear = post_ear.counts_log_prob(test).sum(-1)
    ll_arm = post_arm.counts_log_prob(test).sum()
    ll_van = post_van.counts_log_prob(test[:, None, :]).sum(0)

    def correct(ml, t):
        return torch.gather(t.expand(*ml.shape, t.shape[-1]), -1, ml.long().unsqueeze(-1)).squeeze(-1)
    cor_ear = correct(post_ear.ml_output(seed), test).sum(-1)
    cor_arm = correct(post_arm.ml_output(seed), test).sum()
    cor_van = correct(post_van.ml_output(seed), test[:, None, :]).sum(0)
    return ll_ear, ll_arm, ll_van, cor_ear, cor_arm, cor_van, test.sum()


def evaluation(data, ds_loc_train, ds_loc_test, alphabet, h, ar_func, van_reg, dtype=torch.float64, seed=None):
    """Evaluate a trained BEAR, AR and BMM model (bear_net.py:387-463).

    ds_loc_train : training column to condition on, -1 for none.  h : BEAR concentration parameter.
    van_reg : 1-D array of vanilla-BEAR (BMM) priors.  seed : RNG seed of the argmax tie-breaking
    noise (core.py:69-71); ``seed=-1`` disables the noise (first maximum wins).
    Returns (ll_ear, ll_arm, ll_van, perplexity_ear, perplexity_arm, perplexity_van,
    accuracy_ear, accuracy_arm, accuracy_van) as float64 CPU tensors (``.numpy()`` works).
    """
    table = eng.check_dataset(data)
    if table.A1 != 5:
        return _dense_evaluation(data, ds_loc_train, ds_loc_test, h, ar_func, van_reg, seed)
    head, head_fn = _head_for_eval(ar_func, table)
    hv = float(h.item() if hasattr(h, 'item') else h)
    ll_ear, ll_arm, ll_van, ce, ca, cv, tot = eng.eval_loop(data, ds_loc_train, ds_loc_test, [hv], van_reg,
                                                            head, head_fn, seed)
    return eng.finish_evaluation(ll_ear[0], ll_arm, ll_van, ce[0], ca, cv, tot)


def _dense_evaluation(data, ds_loc_train, ds_loc_test, h, ar_func, van_reg, seed):
    """evaluation() for alphabets the fused kernels do not cover (protein): batches are unpacked on the
    device and run through _evaluation_step (generic distribution kernels)."""
    table = data.table
    use_train = ds_loc_train >= 0
    acc = None
    for r0, n, _ in data.batches():
        for c0 in range(r0, r0 + n, eng.EXPLICIT_CHUNK):
            cn = min(eng.EXPLICIT_CHUNK, r0 + n - c0)
            counts = eng.dense_counts(table, c0, cn)
            batch = [eng.onehot_rows(table, c0, cn), counts[:, ds_loc_test, :]]
            if use_train:
                batch.append(counts[:, ds_loc_train, :])
            with torch.no_grad():
                out = _evaluation_step(batch, h, ar_func, van_reg, table.A1 - 1, use_train, seed=seed)
            acc = list(out) if acc is None else [a + o for a, o in zip(acc, out)]
    acc = [eng.allreduce_sum(a.reshape(-1)).cpu() for a in acc]
    ll_ear, ll_arm, ll_van, ce, ca, cv, tot = acc
    return eng.finish_evaluation(ll_ear[0], ll_arm[0], ll_van, ce[0], ca[0], cv, tot[0])


def h_scan(data, ds_loc_train, ds_loc_test, alphabet, h, ar_func, dtype=torch.float64, seed=None):
    """Evaluate a trained BEAR model at several h values (bear_net.py:465-531).
    Returns (log_likelihood_ear[H], perplexity_ear[H], accuracy_ear[H])."""
    table = eng.check_dataset(data)
    hv = np.asarray(h.cpu() if isinstance(h, torch.Tensor) else h, dtype=np.float64).reshape(-1)
    if table.A1 != 5:
        # protein: the generic (dense-tensor) evaluation, one h at a time; the head is evaluated per h as well --
        # protein tables are short-lag and small, this route is about coverage, not speed
        lls, ces, tot = [], [], None
        for hk in hv:
            use_train = ds_loc_train >= 0
            acc = None
            for r0, n, _ in data.batches():
                for c0 in range(r0, r0 + n, eng.EXPLICIT_CHUNK):
                    cn = min(eng.EXPLICIT_CHUNK, r0 + n - c0)
                    counts = eng.dense_counts(table, c0, cn)
                    batch = [eng.onehot_rows(table, c0, cn), counts[:, ds_loc_test, :]]
                    if use_train:
                        batch.append(counts[:, ds_loc_train, :])
                    with torch.no_grad():
                        out = _evaluation_step(batch, float(hk), ar_func, [1.0], table.A1 - 1, use_train, seed=seed)
                    part = torch.stack([out[0].reshape(()), out[3].reshape(()), out[6].reshape(())])
                    acc = part if acc is None else acc + part
            acc = eng.allreduce_sum(acc).cpu()
            lls.append(acc[0])
            ces.append(acc[1])
            tot = acc[2]
        ll_ear, ce = torch.stack(lls), torch.stack(ces)
        return ll_ear, torch.exp(-ll_ear / tot), ce / tot
    head, head_fn = _head_for_eval(ar_func, table)
    ll_ear, _, _, ce, _, _, tot = eng.eval_loop(data, ds_loc_train, ds_loc_test, hv, [1.0], head, head_fn, seed)
    return ll_ear, torch.exp(-ll_ear / tot), ce / tot
