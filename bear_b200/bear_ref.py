"""Reference-genome BEAR / AR models: the AR function mixes an embedded net with Jukes-Cantor
transition probabilities derived from a reference column of the count table.

Mirrors the reference's ``bear_model/bear_ref.py`` (``_counts_to_probs`` bear_ref.py:9-33,
``_make_ref_ar_func`` :36-69, ``_bear_kmer_counts`` / ``_ar_kmer_counts`` :72-133, ``_create_params`` /
``change_scope_params`` :136-204, ``train`` :262-389, ``evaluation`` :453-539) with the same argument
names and order; params = [h_signed, tau_signed, net_weight_signed, *net params].

Device path: with the stop net (the reference's bear_stop_*.cfg) one fused kernel per training batch
(``bear_ref_train_step``) and one per evaluation batch (``bear_ref_eval_step``, also for the linear net): the
Jukes-Cantor head lives in registers, no f array exists in memory.  Other nets, per training batch: embedded net
g(k) (torch ops / the fused CNN kernels) -> ``bear_ref_head`` (Jukes-Cantor + mixing from the packed reference
column) -> ``bear_dm_train_step_explicit`` (fused DM forward/backward) -> ``bear_ref_head_bwd`` (d tau, d net
weight, d g) -> the net's backward.

Deviation from the reference, on purpose: ``bear_ref._evaluation_step`` reads
``transition_counts_train = batch[3]`` (bear_ref.py:397), which in the mapped tuple
``(onehot, test, train, ref)`` (bear_ref.py:502-507) is the REFERENCE column, not the training one.
No reference test pins that path; this implementation conditions on the training column, as the
docstring of the reference says it should; ``evaluation(..., reference_compatible=True)`` reproduces upstream.
"""
import numpy as np
import torch

from . import _engine as eng
from . import _lib, core
from ._lib import lib, check, ptr

epsilon = 1e-7


def _counts_to_probs(ref_counts, tau, alphabet_size, dtype=torch.float64):
    """Jukes-Cantor transition probabilities from reference counts (bear_ref.py:9-33); ``ref_counts``
    already has epsilon added and stops zeroed (bear_ref.py:332-337)."""
    norm = ref_counts / ref_counts.abs().sum(-1, keepdim=True)
    shape = torch.ones(alphabet_size + 1, dtype=ref_counts.dtype, device=ref_counts.device)
    shape[-1] = 0
    u = (1 / alphabet_size) * shape
    return u + torch.exp(-tau) * (norm - u)


class RefARFunc:
    """ar_func(kmer_seqs, ref_counts) = (nw * net(kmer_seqs) + JC(ref_counts, tau)) / (nw + 1)
    (bear_ref.py:63-68)."""
    kind = 'ref'

    def __init__(self, tau_signed, net_weight_signed, net_func, alphabet_size):
        self.tau_signed, self.net_weight_signed = tau_signed, net_weight_signed
        self.net_func, self.alphabet_size = net_func, alphabet_size

    def __call__(self, kmer_seqs, ref_counts):
        nw = torch.exp(self.net_weight_signed)
        tau = torch.exp(self.tau_signed)
        return (nw * self.net_func(kmer_seqs) + _counts_to_probs(ref_counts, tau, self.alphabet_size)) / (nw + 1)


def _make_ref_ar_func(lag, alphabet_size, make_net_func, af_kwargs, dtype=torch.float64):
    """bear_ref.py:36-69; params = [tau_signed (log 1/30), net_weight_signed (-log 100), *net params]."""
    dev = _lib.device()
    net_weight_signed = torch.tensor(-np.log(100), dtype=dtype, device=dev)
    tau_signed = torch.tensor(np.log(1 / 30), dtype=dtype, device=dev)
    net_func, ar_func_params = make_net_func(lag, alphabet_size, **af_kwargs, dtype=dtype)
    return RefARFunc(tau_signed, net_weight_signed, net_func, alphabet_size), [tau_signed, net_weight_signed] + ar_func_params


def _bear_kmer_counts(kmer_seqs, kmer_total_counts, ref_counts, condition_trans_counts=None, h=None, ar_func=None):
    """bear_ref.py:72-108."""
    dtype, dev = kmer_seqs.dtype, kmer_seqs.device
    if condition_trans_counts is None:
        condition_trans_counts = torch.zeros((), dtype=dtype, device=dev)
    if h is None or ar_func is None:
        h = torch.ones((), dtype=dtype, device=dev)

        def ar_func(x, y):
            return torch.zeros((), dtype=dtype, device=dev)
    concentrations = ar_func(kmer_seqs, ref_counts) / h + condition_trans_counts + epsilon
    return core.tfpDirichletMultinomialPerm(kmer_total_counts, concentrations, name='x')


def _ar_kmer_counts(kmer_seqs, kmer_total_counts, ref_counts, ar_func):
    """bear_ref.py:111-133."""
    return core.tfpMultinomialPerm(kmer_total_counts, ar_func(kmer_seqs, ref_counts) + epsilon, name='x')


def _create_params(lag, alphabet_size, make_ar_func, af_kwargs, dtype=torch.float64):
    """params = [h_signed, tau_signed, net_weight_signed, *net params] (bear_ref.py:136-163)."""
    ar_func, ar_func_params = _make_ref_ar_func(lag, alphabet_size, make_ar_func, af_kwargs, dtype)
    h_signed = torch.zeros((), dtype=dtype, device=_lib.device())
    return [h_signed] + ar_func_params, h_signed, ar_func


def change_scope_params(lag, alphabet_size, make_ar_func, af_kwargs, params, dtype=torch.float64):
    """Unpack a saved parameter list (bear_ref.py:166-204)."""
    from .bear_net import _param_value
    new_params, h_signed, ar_func = _create_params(lag, alphabet_size, make_ar_func, af_kwargs, dtype=dtype)
    if len(params) != len(new_params):
        raise ValueError('expected %d parameters, got %d' % (len(new_params), len(params)))
    for dst, src in zip(new_params, params):
        dst.copy_(_param_value(src, dst.dtype, dst.device).reshape(dst.shape))
    return new_params, h_signed, ar_func


def _net_values(ar_func, table, r0, n):
    """g = net(onehot) [n, A1] for the embedded net, or None for the stop head (handled in-kernel)."""
    net = ar_func.net_func
    if eng.head_kind(net) == 'stop':
        return None
    if eng.fused_cnn_ok(net, table):
        return eng.cnn_forward_autograd(net, table, r0, n) if torch.is_grad_enabled() else eng.cnn_forward(net, table, r0, n)
    g = net(eng.onehot_rows(table, r0, n))
    return g.expand(n, -1) if g.dim() == 1 else g


def _ref_f(ar_func, table, ds_loc_ref, r0, n, g):
    """f for rows [r0, r0+n) through bear_ref_head."""
    gd = None if g is None else g.detach().contiguous()
    f = torch.empty((n, table.A1), dtype=torch.float64, device=_lib.device())
    check(lib.bear_ref_head(table.col_ptr(ds_loc_ref), table.stride, r0, n, ptr(gd), ptr(ar_func.tau_signed),
                            ptr(ar_func.net_weight_signed), ptr(f), _lib.stream()))
    return f, gd


def train(data, num_kmers, epochs, ds_loc, ds_loc_ref, alphabet, lag, make_ar_func, af_kwargs,
          learning_rate, optimizer_name, train_ar, acc_steps=1,
          params_restart=None, writer=None, loss_save=None, dtype=torch.float64):
    """Train a BEAR or AR model based on reference transition counts (bear_ref.py:262-389).
    Same arguments as ``bear_net.train`` plus ``ds_loc_ref``, the reference column."""
    table = eng.check_dataset(data)
    if table.A1 != 5:
        raise ValueError('bear_ref kernels cover the DNA / RNA alphabets')
    alphabet_size = len(core.alphabets_tf[alphabet]) - 1
    if params_restart is None:
        params, h_signed, ar_func = _create_params(lag, alphabet_size, make_ar_func, af_kwargs)
    else:
        params, h_signed, ar_func = change_scope_params(lag, alphabet_size, make_ar_func, af_kwargs, params_restart)
    fp = eng.FlatParams(params)
    ws = eng.workspace(table, fp.total)
    for p in params[3:]:
        p.requires_grad_(True)

    stop_net = table.A1 == 5 and eng.head_kind(ar_func.net_func) == 'stop'

    def step_fn(r0, n, scale):
        if stop_net:
            # stop net (bear_stop_bear.cfg / bear_stop_ar.cfg): head, loss and the gradients of [h, tau, net weight] in ONE
            # kernel over the data and reference columns; fp.grad = [loss, d h, d tau_signed, d net_weight_signed]
            check(lib.bear_ref_train_step(table.col_ptr(ds_loc), table.col_ptr(ds_loc_ref), table.stride, r0, n, ptr(h_signed),
                                          ptr(ar_func.tau_signed), ptr(ar_func.net_weight_signed), scale, int(train_ar),
                                          ptr(fp.grad), None, ptr(ws), _lib.stream()))
            return
        for c0 in range(r0, r0 + n, eng.EXPLICIT_CHUNK):
            cn = min(eng.EXPLICIT_CHUNK, r0 + n - c0)
            with torch.enable_grad():
                g = _net_values(ar_func, table, c0, cn)
            f, gd = _ref_f(ar_func, table, ds_loc_ref, c0, cn, g)
            gf = torch.empty_like(f)
            check(lib.bear_dm_train_step_explicit(table.col_ptr(ds_loc), table.stride, c0, cn, ptr(f), ptr(h_signed),
                                                  scale, int(train_ar), ptr(fp.grad), ptr(gf), None, ptr(ws),
                                                  _lib.stream()))
            gg = torch.empty_like(gd) if gd is not None else None
            # fp.grad = [loss, d h, d tau_signed, d net_weight_signed, ...]
            check(lib.bear_ref_head_bwd(table.col_ptr(ds_loc_ref), table.stride, c0, cn, ptr(gd), ptr(ar_func.tau_signed),
                                        ptr(ar_func.net_weight_signed), ptr(gf), ptr(gg), ptr(fp.grad[2:4]), ptr(ws),
                                        _lib.stream()))
            if g is not None and g.requires_grad:
                g.backward(gg)
        for i, t in enumerate(fp.tensors):
            if t.grad is not None:
                fp.grad_view(i).add_(t.grad)
                t.grad = None

    eng.train_loop(data, num_kmers, ds_loc, train_ar, acc_steps, fp, optimizer_name, learning_rate, step_fn,
                   writer=writer, loss_save=loss_save, graph_safe=stop_net)
    for p in params:
        p.requires_grad_(False)
    return params, h_signed, ar_func


def evaluation(data, ds_loc_train, ds_loc_test, ds_loc_ref, alphabet, h, ar_func, van_reg, dtype=torch.float64,
               seed=None, reference_compatible=False):
    """Evaluate a trained reference-based BEAR, AR and BMM model (bear_ref.py:453-539); returns the same
    9-tuple as ``bear_net.evaluation``.

    ``reference_compatible=True`` reproduces the upstream quirk described in the module docstring: with
    ``ds_loc_train >= 0`` the BEAR and BMM posteriors are conditioned on the MAPPED REFERENCE column
    ``(ref + eps) * not_stop`` (bear_ref.py:397 reads batch[3] of the tuple built at :502-507) instead of the training
    column -- for comparing numbers with runs of the reference.  That route goes through the generic (dense-tensor)
    evaluation, not the fused kernel."""
    table = eng.check_dataset(data)

    def f_fn(c0, cn):
        return _ref_f(ar_func, table, ds_loc_ref, c0, cn, _net_values(ar_func, table, c0, cn))[0]

    if reference_compatible and ds_loc_train >= 0:
        from . import bear_net
        not_stop = torch.ones(table.A1, dtype=torch.float64, device=_lib.device())
        not_stop[-1] = 0.0
        acc = None
        for r0, n, _ in data.batches():
            for c0 in range(r0, r0 + n, eng.EXPLICIT_CHUNK):
                cn = min(eng.EXPLICIT_CHUNK, r0 + n - c0)
                counts = eng.dense_counts(table, c0, cn)
                f = f_fn(c0, cn)
                batch = [eng.onehot_rows(table, c0, cn), counts[:, ds_loc_test, :],
                         (counts[:, ds_loc_ref, :] + epsilon) * not_stop]
                with torch.no_grad():
                    out = bear_net._evaluation_step(batch, h, lambda x, f=f: f, van_reg, table.A1 - 1, True, seed=seed)
                acc = list(out) if acc is None else [a + o for a, o in zip(acc, out)]
        acc = [eng.allreduce_sum(a.reshape(-1)).cpu() for a in acc]
        ll_ear, ll_arm, ll_van, ce, ca, cv, tot = acc
        return eng.finish_evaluation(ll_ear[0], ll_arm[0], ll_van, ce[0], ca[0], cv, tot[0])

    hv = float(h.item() if hasattr(h, 'item') else h)
    # stop and linear nets on DNA / RNA tables: ONE fused kernel per batch (bear_ref_eval_step: Jukes-Cantor head in
    # registers); other nets: f is materialised by torch ops / the CNN kernel and read back by bear_eval_step
    ref = None
    kind = eng.head_kind(ar_func.net_func)
    if table.A1 == 5 and kind == 'stop':
        ref = (ds_loc_ref, _lib.HEAD_STOP, None, ar_func.tau_signed, ar_func.net_weight_signed)
    elif eng.fused_linear_ok(ar_func.net_func, table):
        ref = (ds_loc_ref, _lib.HEAD_LINEAR, ar_func.net_func.params[0].detach().contiguous(), ar_func.tau_signed,
               ar_func.net_weight_signed)
    ll_ear, ll_arm, ll_van, ce, ca, cv, tot = eng.eval_loop(data, ds_loc_train, ds_loc_test, [hv], van_reg,
                                                            _lib.HEAD_EXPLICIT, eng.explicit_head_ptr_fn(f_fn), seed, ref=ref)
    return eng.finish_evaluation(ll_ear[0], ll_arm, ll_van, ce[0], ca, cv, tot)
