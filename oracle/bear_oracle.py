"""CPU oracle for the BEAR hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` leg may import this module.  Nothing under ``bear_b200/``
imports it; the product path fails loudly when the CUDA library is missing.

What it is: an op-for-op float64 restatement (torch-CPU autograd + numpy/scipy)
of the reference's TensorFlow graph for the Dirichlet-multinomial marginal
log-likelihood path.  Every function cites the reference file:line it follows
(paths relative to /root/reference/bear_model/).

Third-party arithmetic that is NOT under /root/reference and is restated here
from its published algorithm:
  * tensorflow_probability==0.11.1 (requirements.txt:11)
      DirichletMultinomial.log_prob(c) = lbeta(conc + c) - lbeta(conc) + log_combinations(n, c)
      Multinomial.log_prob(c)          = sum(multiply_no_nan(log(probs), c)) + log_combinations(n, c)
      math.log_combinations(n, c)      = lgamma(n + 1) - sum(lgamma(c + 1))
  * tensorflow (unpinned, requirements.txt:10)
      math.lbeta(x) = sum(lgamma(x)) - lgamma(sum(x)); nn.softmax; nn.moments
      (biased variance); nn.elu; nn.conv1d VALID; keras.backend.epsilon() = 1e-7;
      keras.optimizers.Adam (OptimizerV2):
          lr_t = lr*sqrt(1-b2^t)/(1-b1^t); m,v EMA; theta -= lr_t*m/(sqrt(v)+1e-7)

Pinning status: the reference itself cannot be imported in this image
(tensorflow / tensorflow_probability / tensorflow_io are absent), so the oracle
is pinned against the reference's OWN known-answer tests and golden vectors
(tests/test_core.py:23-26,59-60; tests/test_dataloader.py:25-32,42-49;
tests/test_run.py:26-30; tests/test_var_prob.py:66-78,152-173;
docs/usage.rst:255-265) -- see tests/test_oracle_pinned.py.  Gradient parity is
unpinned by the reference (it has no gradient tests); the oracle's gradients are
torch autograd of the restated graph, cross-checked by finite differences.
"""
import math

import numpy as np
import torch

EPS = 1e-7  # tf.keras.backend.epsilon(); core.py:8, bear_net.py:4

# core.py:142-153.  Input symbols end with the start token '[', count columns end
# with the stop token ']'.
ALPHABETS_IN = {
    'prot': list('ARNDCEQGHILKMFPSTWYV') + ['['],
    'dna': list('ACGT') + ['['],
    'rna': list('ACGU') + ['['],
}
ALPHABETS_OUT = {
    'prot': list('ARNDCEQGHILKMFPSTWYV') + [']'],
    'dna': list('ACGT') + [']'],
    'rna': list('ACGU') + [']'],
}


# ----------------------------------------------------------------------------
# data
# ----------------------------------------------------------------------------
def one_hot(kmers, alphabet='dna', dtype=torch.float64):
    """core.py:156-174 tf_one_hot: bytes_split -> equal(alphabet) -> cast.
    Unknown characters give an all-zero row."""
    alph = ALPHABETS_IN[alphabet]
    kmers = [k.decode() if isinstance(k, bytes) else str(k) for k in kmers]
    lag = len(kmers[0]) if kmers else 0
    out = torch.zeros(len(kmers), lag, len(alph), dtype=dtype)
    for i, k in enumerate(kmers):
        for j, ch in enumerate(k):
            if ch in alph:
                out[i, j, alph.index(ch)] = 1.0
    return out


def one_hot_bytes(kmers, alphabet='dna', dtype=torch.float64):
    """core.py:156-174 on a numpy array of equal-length byte strings (dtype 'S<lag>'), vectorised the way the TF
    graph is: bytes_split -> equal against the alphabet (broadcast) -> cast.  Same values as ``one_hot``."""
    kmers = np.asarray(kmers)
    lag = kmers.dtype.itemsize
    sym = kmers.view(np.uint8).reshape(len(kmers), lag)
    alph = np.frombuffer(''.join(ALPHABETS_IN[alphabet]).encode(), dtype=np.uint8)
    return torch.from_numpy(sym[:, :, None] == alph[None, None, :]).to(dtype)


def read_tsv(path, num_ds, alphabet='dna', header=False):
    """dataloader.py:6-50 dense TSV ``kmer \\t [[g0...],[g1...],...]``.
    Returns (list of kmer str, float64 ndarray [K, num_ds, A+1])."""
    import json
    a1 = len(ALPHABETS_IN[alphabet])
    kmers, counts = [], []
    with open(path) as fh:
        for ln, line in enumerate(fh):
            if header and ln == 0:
                continue
            line = line.rstrip('\n')
            if not line:
                continue
            k, mat = line.split('\t')
            kmers.append(k)
            counts.append(json.loads(mat))
    counts = np.asarray(counts, dtype=np.float64).reshape(len(kmers), num_ds, a1)
    return kmers, counts


def read_sparse(path, num_ds, alphabet='dna', header=True):
    """dataloader.py:52-109 sparse ``kmer; [[g,b],...]; [v,...]`` (';'-separated)."""
    import json
    a1 = len(ALPHABETS_IN[alphabet])
    kmers, rows = [], []
    with open(path) as fh:
        for ln, line in enumerate(fh):
            if header and ln == 0:
                continue
            line = line.strip()
            if not line:
                continue
            k, pos, val = [s.strip() for s in line.split(';')]
            dense = np.zeros((num_ds, a1))
            for (g, b), v in zip(json.loads(pos), json.loads(val)):
                dense[g, b] += v
            kmers.append(k)
            rows.append(dense)
    return kmers, np.asarray(rows, dtype=np.float64).reshape(len(kmers), num_ds, a1)


# ----------------------------------------------------------------------------
# distributions (core.py)
# ----------------------------------------------------------------------------
def _t(x, dtype=torch.float64):
    return x if isinstance(x, torch.Tensor) else torch.as_tensor(np.array(x, dtype=np.float64), dtype=dtype)


def lbeta(x):
    """tf.math.lbeta over the last axis."""
    return torch.lgamma(x).sum(-1) - torch.lgamma(x.sum(-1))


def log_combinations(n, counts):
    """tfp.math.log_combinations (TFP 0.11.1)."""
    return torch.lgamma(n + 1.0) - torch.lgamma(counts + 1.0).sum(-1)


def dm_counts_log_prob(total_count, concentration, value, with_cancelling_term=True):
    """core.py:60-62,73-74: DirichletMultinomial.log_prob(value) - log_combinations.
    ``with_cancelling_term`` keeps the reference's add-then-subtract of the
    multinomial coefficient (pure rounding noise)."""
    total_count, concentration, value = _t(total_count), _t(concentration), _t(value)
    ordered = lbeta(concentration + value) - lbeta(concentration + 0.0 * value)
    if with_cancelling_term:
        lc = log_combinations(total_count, value)
        return (ordered + lc) - lc
    return ordered


def mn_counts_log_prob(total_count, probs, value, with_cancelling_term=True):
    """core.py:125-127,138-139: Multinomial.log_prob(value) - log_combinations;
    multiply_no_nan(log p, c) gives 0 where c == 0 even if log p = -inf."""
    total_count, probs, value = _t(total_count), _t(probs), _t(value)
    logp = torch.log(probs)
    prod = torch.where(value == 0, torch.zeros_like(value * logp), value * logp)
    out = prod.sum(-1)
    if with_cancelling_term:
        lc = log_combinations(total_count, value)
        return (out + lc) - lc
    return out


def ml_output_noiseless(conc):
    """core.py:69-71,134-136 without the tie-breaking noise (argmax)."""
    return torch.argmax(_t(conc), dim=-1)


# ----------------------------------------------------------------------------
# AR heads (ar_funcs.py)
# ----------------------------------------------------------------------------
def normalize_layer(x):
    """ar_funcs.py:5-20: (x-mean)/sqrt(biased var + 1e-5) over the last axis."""
    mean = x.mean(-1, keepdim=True)
    var = ((x - mean) ** 2).mean(-1, keepdim=True)
    return (x - mean) / torch.sqrt(var + 1e-5)


def init_linear(lag, alphabet_size, gen):
    """ar_funcs.py:41-42: 0.05 * l2_normalize(N(0,1), axis=1)."""
    a1 = alphabet_size + 1
    mat = torch.randn(lag, a1, a1, dtype=torch.float64, generator=gen)
    mat = 0.05 * mat / torch.sqrt((mat ** 2).sum(1, keepdim=True).clamp_min(1e-12))
    return [mat]


def ar_linear(onehot, params):
    """ar_funcs.py:44-45: softmax(einsum('...jk,jkl->...l'))."""
    (mat,) = params
    return torch.softmax(torch.einsum('...jk,jkl->...l', onehot, mat), dim=-1)


def init_cnn(lag, alphabet_size, gen, filter_width=8, num_filters=30, kmer_layer1_width=16):
    """ar_funcs.py:72-89. Returned in the reference's param order (ar_funcs.py:98-99):
    [filters, int0, W1, int1, W2, int2, scale0, scale1]."""
    W, F, H1 = int(filter_width), int(num_filters), int(kmer_layer1_width)
    a1 = alphabet_size + 1
    P = lag - W + 1
    f64 = torch.float64

    def l2n(x, dims):
        return x / torch.sqrt((x ** 2).sum(dims, keepdim=True).clamp_min(1e-12))
    filters = l2n(torch.randn(W, a1, F, dtype=f64, generator=gen), (0, 1))
    int0 = torch.ones(P, F, dtype=f64)
    scale0 = torch.ones(P, F, dtype=f64)
    W1 = l2n(torch.randn(P, F, H1, dtype=f64, generator=gen), (0,))
    int1 = torch.ones(H1, dtype=f64)
    scale1 = torch.ones(H1, dtype=f64)
    W2 = 0.05 * l2n(torch.randn(H1, a1, dtype=f64, generator=gen), (0,))
    int2 = torch.zeros(a1, dtype=f64)
    return [filters, int0, W1, int1, W2, int2, scale0, scale1]


def ar_cnn(onehot, params):
    """ar_funcs.py:91-97."""
    filters, int0, W1, int1, W2, int2, scale0, scale1 = params
    W = filters.shape[0]
    L = onehot.shape[-2]
    P = L - W + 1
    # conv1d VALID, stride 1: out[..., p, f] = sum_{w,a} x[..., p+w, a] * filters[w, a, f]
    win = torch.stack([onehot[..., p:p + W, :] for p in range(P)], dim=-3)  # [..., P, W, A1]
    conv = torch.einsum('...pwa,waf->...pf', win, filters)
    x0 = scale0 * normalize_layer(conv) + int0
    x1 = scale1 * normalize_layer(torch.einsum('...pf,pfh->...h', torch.nn.functional.elu(x0), W1)) + int1
    x2 = torch.nn.functional.elu(x1) @ W2 + int2
    return torch.softmax(x2, dim=-1)


def ar_stop(onehot, alphabet_size):
    """ar_funcs.py:121-126: constant [0,...,0,1]."""
    v = torch.zeros(alphabet_size + 1, dtype=torch.float64)
    v[-1] = 1.0
    return v


AR_FUNCS = {'linear': ar_linear, 'cnn': ar_cnn}


# ----------------------------------------------------------------------------
# bear_ref head (bear_ref.py:9-69)
# ----------------------------------------------------------------------------
def ref_counts_map(counts_ref, alphabet_size):
    """bear_ref.py:332-337: (ref + eps) * not_stop."""
    not_stop = torch.ones(alphabet_size + 1, dtype=torch.float64)
    not_stop[-1] = 0.0
    return (_t(counts_ref) + EPS) * not_stop


def counts_to_probs(ref_counts, tau, alphabet_size):
    """bear_ref.py:9-33: L1 normalise, Jukes-Cantor mix with the uniform-no-stop vector."""
    norm = ref_counts / ref_counts.abs().sum(-1, keepdim=True)
    shape = torch.ones(alphabet_size + 1, dtype=torch.float64)
    shape[-1] = 0.0
    u = (1.0 / alphabet_size) * shape
    return u + torch.exp(-tau) * (norm - u)


def ar_ref(onehot, ref_counts, tau_signed, nw_signed, net_func, alphabet_size):
    """bear_ref.py:63-68."""
    nw = torch.exp(nw_signed)
    tau = torch.exp(tau_signed)
    return (nw * net_func(onehot) + counts_to_probs(ref_counts, tau, alphabet_size)) / (nw + 1.0)


# ----------------------------------------------------------------------------
# train step (bear_net.py:146-197, bear_ref.py:207-259)
# ----------------------------------------------------------------------------
def train_loss(onehot, counts, h_signed, ar_func, num_kmers, train_ar, with_cancelling_term=True):
    """Returns the scalar loss = -(num_kmers / B) * sum_k ll_k (bear_net.py:176-191)
    and the per-k-mer log-likelihoods."""
    B = onehot.shape[0]
    total = counts.sum(-1)
    f = ar_func(onehot)
    if train_ar:
        ll = mn_counts_log_prob(total, f + EPS, counts, with_cancelling_term)  # bear_net.py:68
    else:
        conc = f / torch.exp(h_signed) + 0.0 + EPS                               # bear_net.py:43
        ll = dm_counts_log_prob(total, conc, counts, with_cancelling_term)
    loss = -(num_kmers / B) * ll.sum()
    return loss, ll


def train_step_grads(onehot, counts, h_signed, params, head, num_kmers, train_ar,
                     with_cancelling_term=True):
    """loss and d loss / d [h_signed] + params via autograd (bear_net.py:193).
    A parameter with no gradient path (h_signed when train_ar) gets zeros, which
    is what acc_grads holds in the reference (bear_net.py:194-196)."""
    h = h_signed.clone().requires_grad_(True)
    ps = [p.clone().requires_grad_(True) for p in params]
    loss, ll = train_loss(onehot, counts, h, lambda x: AR_FUNCS[head](x, ps), num_kmers,
                          train_ar, with_cancelling_term)
    grads = torch.autograd.grad(loss, [h] + ps, allow_unused=True)
    grads = [torch.zeros_like(p) if g is None else g for g, p in zip(grads, [h] + ps)]
    return loss.detach(), ll.detach(), grads


class KerasAdam:
    """tf.keras.optimizers.Adam (OptimizerV2 defaults) restated; used by
    bear_net.py:264-265,277-282."""

    def __init__(self, params, learning_rate, beta_1=0.9, beta_2=0.999, epsilon=1e-7):
        self.lr, self.b1, self.b2, self.eps = learning_rate, beta_1, beta_2, epsilon
        self.m = [torch.zeros_like(p) for p in params]
        self.v = [torch.zeros_like(p) for p in params]
        self.t = 0

    def apply(self, params, grads):
        self.t += 1
        lr_t = self.lr * math.sqrt(1.0 - self.b2 ** self.t) / (1.0 - self.b1 ** self.t)
        for p, g, m, v in zip(params, grads, self.m, self.v):
            m.mul_(self.b1).add_(g, alpha=1.0 - self.b1)
            v.mul_(self.b2).addcmul_(g, g, value=1.0 - self.b2)
            p.sub_(lr_t * m / (v.sqrt() + self.eps))


def train(batches, num_kmers, head, params, h_signed, learning_rate, train_ar, acc_steps=1,
          loss_save=None):
    """bear_net.py:293-315 training loop on pre-encoded batches [(onehot, counts), ...]."""
    params = [p.clone() for p in params]
    h_signed = h_signed.clone()
    allp = [h_signed] + params
    opt = KerasAdam(allp, learning_rate)
    acc = [torch.zeros_like(p) for p in allp]
    loss_acc, step = 0.0, 1
    for onehot, counts in batches:
        loss, _, grads = train_step_grads(onehot, counts, h_signed, params, head, num_kmers, train_ar)
        loss_acc += float(loss)
        for a, g in zip(acc, grads):
            a.add_(g)
        if step % acc_steps == 0:
            if loss_save is not None:
                loss_save.append(-loss_acc / acc_steps)
            opt.apply(allp, acc)
            for a in acc:
                a.zero_()
            loss_acc = 0.0
        step += 1
    return params, h_signed


# ----------------------------------------------------------------------------
# evaluation (bear_net.py:323-371,459-463; h_scan :516-531)
# ----------------------------------------------------------------------------
def evaluation_step(onehot, test, train, h, f, van_reg):
    """One batch of bear_net._evaluation_step without the argmax noise.
    ``h`` is a scalar or [H] tensor (h_scan broadcasts it as [H,1,1], bear_net.py:522-523);
    ``f`` = ar_func(onehot) [B, A+1].  Returns ll_ear ([] or [H]), ll_arm, ll_van [V],
    correct_ear, correct_arm, correct_van [V], total_len."""
    test = _t(test)
    van_reg = _t(van_reg)
    a1 = test.shape[-1]
    total = test.sum(-1)
    h = _t(h)
    scan = h.dim() > 0
    if train is not None:
        train = _t(train)
        van_cond = train[:, None, :] + van_reg[:, None]                     # bear_net.py:328
        cond = train
    else:
        van_cond = van_reg[:, None] * torch.ones(1, a1, dtype=torch.float64)  # bear_net.py:331
        cond = torch.zeros((), dtype=torch.float64)
    hh = h.reshape(-1, 1, 1) if scan else h
    conc_ear = f / hh + cond + EPS                                          # bear_net.py:43
    ll_ear = dm_counts_log_prob(total, conc_ear, test).sum(-1)
    p_arm = f + EPS
    ll_arm = mn_counts_log_prob(total, p_arm, test).sum()
    conc_van = torch.zeros(()) + van_cond + EPS                             # ar_func==0, h==1
    if conc_van.dim() == 2:
        conc_van = conc_van[None].expand(test.shape[0], -1, -1)
    ll_van = dm_counts_log_prob(total[:, None], conc_van, test[:, None, :]).sum(0)

    def correct(conc, t):
        idx = torch.argmax(conc, dim=-1, keepdim=True)
        return torch.gather(t.expand(conc.shape), -1, idx).squeeze(-1)
    cor_ear = correct(conc_ear, test).sum(-1)
    cor_arm = correct(p_arm, test).sum()
    cor_van = correct(conc_van, test[:, None, :]).sum(0)
    return ll_ear, ll_arm, ll_van, cor_ear, cor_arm, cor_van, test.sum()


def evaluation(batches, h, van_reg):
    """bear_net.py:439-463 over [(onehot, f, test, train_or_None), ...]."""
    acc = None
    for onehot, f, test, train in batches:
        out = evaluation_step(onehot, test, train, h, f, van_reg)
        acc = list(out) if acc is None else [a + o for a, o in zip(acc, out)]
    ll_ear, ll_arm, ll_van, ce, ca, cv, tot = acc
    return (ll_ear, ll_arm, ll_van,
            torch.exp(-ll_ear / tot), torch.exp(-ll_arm / tot), torch.exp(-ll_van / tot),
            ce / tot, ca / tot, cv / tot)


def bmm_likelihood(counts, alpha):
    """dataloader.py:111-113: sum_k lbeta(c + a) - lbeta(0*c + a) -> [G, V]."""
    counts, alpha = _t(counts), _t(alpha)
    x = counts[..., None, :] + alpha[:, None]
    x0 = 0.0 * counts[..., None, :] + alpha[:, None]
    return lbeta(x).sum(0) - lbeta(x0).sum(0)


# ----------------------------------------------------------------------------
# get_pdf numeric core (get_var_probs.py:132-175) and the sampler (log_gamma.py)
# ----------------------------------------------------------------------------
def get_pdf_concs(counts_train, ar_vals, h, vans, get_map):
    """get_var_probs.py:132-153 -> concs [num_models, K, A+1]."""
    counts_train = np.asarray(counts_train, dtype=np.float64)
    K, a1 = counts_train.shape
    alpha = None
    if len(vans) > 0:
        alpha = np.array(vans)[:, None, None] * np.ones([K, a1])[None, ...]
    if ar_vals is not None:
        dec = ar_vals[None, :, :] / np.asarray(h)[:, None, None]
        alpha = dec if alpha is None else np.concatenate([dec, alpha], axis=0)
    concs = alpha + counts_train[None, :, :]
    if ar_vals is not None and get_map:
        concs = np.concatenate([ar_vals[None, ...], concs], axis=0)
    return concs


def get_pdf_map(concs):
    """get_var_probs.py:172."""
    return np.log(concs / np.sum(concs, axis=-1)[..., None])


def get_pdf_marg(concs_kmc, counts):
    """get_var_probs.py:165-169 for concs [K, M, A+1], counts [K, A+1] -> [M]."""
    from scipy.special import loggamma
    lp = (loggamma(np.sum(concs_kmc, axis=-1)) - np.sum(loggamma(concs_kmc), axis=-1)
          - loggamma(np.sum(concs_kmc + counts[:, None, :], axis=-1))
          + np.sum(loggamma(concs_kmc + counts[:, None, :]), axis=-1))
    return np.sum(lp, axis=0)


def log_gamma_sample(concs, size, rng):
    """log_gamma.py:17-76 restated with a numpy Generator (same proposal / accept rule)."""
    from scipy.special import gammaln
    import scipy.stats as st
    concs = np.asarray(concs, dtype=np.float64)
    shape = np.r_[size, np.shape(concs)].astype(int)
    concs = np.tile(concs, np.r_[size, np.ones(len(np.shape(concs)))].astype(int)).flatten()
    n = len(concs)
    draws = np.empty(n)
    big = concs >= 1
    draws[big] = np.log(rng.standard_gamma(concs[big]))
    remain = np.flatnonzero(~big)
    with np.errstate(divide='ignore'):
        log_prob_neg = np.log(st.gamma.cdf(1, concs))
    log_ms = np.maximum(0, -(log_prob_neg + gammaln(concs) + np.log(concs)))
    while remain.size:
        c = concs[remain]
        x_gam = rng.standard_gamma(c)
        pos = x_gam > 1
        acc = np.zeros(remain.size, dtype=bool)
        val = np.empty(remain.size)
        u = rng.uniform(size=remain.size)
        val[pos] = np.log(x_gam[pos])
        acc[pos] = u[pos] < np.exp(-log_ms[remain][pos])
        x_neg = -rng.standard_gamma(np.ones((~pos).sum())) / c[~pos]
        ratio = np.exp(-np.exp(x_neg) - gammaln(c[~pos]) - np.log(c[~pos])
                       - log_prob_neg[remain][~pos] - log_ms[remain][~pos])
        val[~pos] = x_neg
        acc[~pos] = u[~pos] < ratio
        draws[remain[acc]] = val[acc]
        remain = remain[~acc]
    return draws.reshape(shape)
