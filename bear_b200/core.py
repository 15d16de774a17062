"""Order-aware Dirichlet-multinomial / multinomial distributions, alphabets and one-hot encoding.

Mirrors the reference's ``bear_model/core.py`` (``tfpDirichletMultinomialPerm`` core.py:11-74,
``tfpMultinomialPerm`` core.py:77-139, ``alphabets_tf`` / ``alphabets_en`` core.py:142-153,
``tf_one_hot`` core.py:156-174) on torch CUDA tensors.  The arithmetic runs in libbear_b200's
generic dense kernels; gradients flow to ``concentration`` / ``probs`` through the analytic
digamma backward (the reference gets them from a GradientTape, bear_net.py:193).
"""
import numpy as np
import torch

from . import _lib
from ._lib import lib, check, ptr

epsilon = 1e-7     # tf.keras.backend.epsilon(), core.py:8

alphabets_tf = {
    'prot': np.array([b'A', b'R', b'N', b'D', b'C', b'E', b'Q', b'G', b'H', b'I', b'L',
                      b'K', b'M', b'F', b'P', b'S', b'T', b'W', b'Y', b'V', b'[']),
    'dna': np.array([b'A', b'C', b'G', b'T', b'[']),
    'rna': np.array([b'A', b'C', b'G', b'U', b'['])}

alphabets_en = {
    'prot': np.array(['A', 'R', 'N', 'D', 'C', 'E', 'Q', 'G', 'H', 'I', 'L',
                      'K', 'M', 'F', 'P', 'S', 'T', 'W', 'Y', 'V', ']']),
    'dna': np.array(['A', 'C', 'G', 'T', ']']),
    'rna': np.array(['A', 'C', 'G', 'U', ']'])}


def _as_f64_cuda(x):
    if isinstance(x, torch.Tensor):
        t = x
    else:
        t = torch.as_tensor(np.asarray(x, dtype=np.float64))
    if not t.is_cuda:
        t = t.to(_lib.device())
    return t.to(torch.float64)


def _broadcast_rows(param, value_shape):
    """Returns (param_2d contiguous [rows, A1], rows) such that flattened value row i uses
    param_2d[i % rows]; falls back to materialising the broadcast when the pattern is not periodic."""
    A1 = value_shape[-1]
    lead = tuple(value_shape[:-1])
    p = param
    while p.dim() > 1 and p.shape[0] == 1:
        p = p[0]
    pshape = tuple(p.shape[:-1])
    if p.shape[-1] == A1 and len(pshape) <= len(lead) and pshape == lead[len(lead) - len(pshape):]:
        rows = int(np.prod(pshape)) if pshape else 1
        return p.reshape(rows, A1).contiguous(), rows
    full = param.expand(*lead, A1).contiguous().reshape(-1, A1)
    return full, full.shape[0]


class _DMLogProb(torch.autograd.Function):
    @staticmethod
    def forward(ctx, conc, value):
        lead = torch.broadcast_shapes(conc.shape[:-1], value.shape[:-1])
        A1 = value.shape[-1]
        v = value.expand(*lead, A1).contiguous().reshape(-1, A1)
        c2, rows = _broadcast_rows(conc, (*lead, A1))
        out = torch.empty(v.shape[0], dtype=torch.float64, device=v.device)
        check(lib.bear_dm_logprob(ptr(c2), rows, ptr(v), v.shape[0], A1, ptr(out), _lib.stream()))
        ctx.save_for_backward(c2, v)
        ctx.meta = (rows, lead, A1, conc.shape)
        return out.reshape(lead)

    @staticmethod
    def backward(ctx, gout):
        c2, v = ctx.saved_tensors
        rows, lead, A1, cshape = ctx.meta
        g = gout.contiguous().reshape(-1)
        gc = torch.empty_like(v)
        check(lib.bear_dm_logprob_bwd(ptr(c2), rows, ptr(v), v.shape[0], A1, ptr(g), ptr(gc), _lib.stream()))
        return gc.reshape(*lead, A1).sum_to_size(cshape), None


class _MNLogProb(torch.autograd.Function):
    @staticmethod
    def forward(ctx, probs, value):
        lead = torch.broadcast_shapes(probs.shape[:-1], value.shape[:-1])
        A1 = value.shape[-1]
        v = value.expand(*lead, A1).contiguous().reshape(-1, A1)
        p2, rows = _broadcast_rows(probs, (*lead, A1))
        out = torch.empty(v.shape[0], dtype=torch.float64, device=v.device)
        check(lib.bear_mn_logprob(ptr(p2), rows, ptr(v), v.shape[0], A1, ptr(out), _lib.stream()))
        ctx.save_for_backward(p2, v)
        ctx.meta = (rows, lead, A1, probs.shape)
        return out.reshape(lead)

    @staticmethod
    def backward(ctx, gout):
        p2, v = ctx.saved_tensors
        rows, lead, A1, pshape = ctx.meta
        g = gout.contiguous().reshape(-1)
        gp = torch.empty_like(v)
        check(lib.bear_mn_logprob_bwd(ptr(p2), rows, ptr(v), v.shape[0], A1, ptr(g), ptr(gp), _lib.stream()))
        return gp.reshape(*lead, A1).sum_to_size(pshape), None


def _ml_output(x, sigma, seed):
    A1 = x.shape[-1]
    flat = x.contiguous().reshape(-1, A1)
    out = torch.empty(flat.shape[0], dtype=torch.float64, device=flat.device)
    if seed is None:
        seed = int(torch.randint(0, 2 ** 62, (1,)).item())
    check(lib.bear_ml_output(ptr(flat), flat.shape[0], A1, sigma, seed, ptr(out), _lib.stream()))
    return out.reshape(x.shape[:-1])


class tfpDirichletMultinomialPerm:
    """Dirichlet-multinomial over transition counts observed in a particular order (core.py:11-74).

    total_count : [A1..An]; concentration : [Am..An, alphabet_size+1] (broadcasts from the right).
    ``counts_log_prob(value)`` = DirichletMultinomial.log_prob - log_combinations, i.e.
    sum_b lgamma(conc_b + c_b) - lgamma(conc_b) - [lgamma(S + N) - lgamma(S)]; total_count only
    enters the cancelling combinatorial term and is unused.
    """

    def __init__(self, total_count, concentration, validate_args=False, allow_nan_stats=True,
                 name='DirichletMultinomialPerm'):
        self.total_count = _as_f64_cuda(total_count)
        self.concentration = _as_f64_cuda(concentration)
        self.alphabet_size = int(self.concentration.shape[-1]) - 1
        self.dtype = self.concentration.dtype
        self.name = name

    def _sample_n(self, n, seed=None, dummy=True):
        """Zeros of shape [n, *total_count.shape, alphabet_size+1] (core.py:64-67)."""
        return torch.zeros((n, *self.total_count.shape, self.alphabet_size + 1), dtype=self.dtype,
                           device=self.concentration.device)

    def ml_output(self, seed=None):
        """argmax of concentration + 100 eps N(0,1) (core.py:69-71); ``seed=-1`` disables the noise."""
        conc = self.concentration
        lead = torch.broadcast_shapes(conc.shape[:-1], self.total_count.shape)
        conc = conc.expand(*lead, conc.shape[-1])
        return _ml_output(conc, 100 * epsilon, seed)

    def counts_log_prob(self, value):
        return _DMLogProb.apply(self.concentration, _as_f64_cuda(value))


class tfpMultinomialPerm:
    """Multinomial over ordered transition counts (core.py:77-139): sum_b c_b log p_b with
    0 * log 0 = 0; probs are used as given (not renormalised)."""

    def __init__(self, total_count, probs, validate_args=False, allow_nan_stats=True,
                 name='DirichletMultinomialPerm'):
        self.total_count = _as_f64_cuda(total_count)
        self.probs = _as_f64_cuda(probs)
        self.alphabet_size = int(self.probs.shape[-1]) - 1
        self.dtype = self.probs.dtype
        self.name = name

    def _sample_n(self, n, seed=None):
        return torch.zeros((n, *self.total_count.shape, self.alphabet_size + 1), dtype=self.dtype,
                           device=self.probs.device)

    def ml_output(self, seed=None):
        """argmax of probs + eps N(0,1) (core.py:134-136)."""
        p = self.probs
        lead = torch.broadcast_shapes(p.shape[:-1], self.total_count.shape)
        return _ml_output(p.expand(*lead, p.shape[-1]), epsilon, seed)

    def counts_log_prob(self, value):
        return _MNLogProb.apply(self.probs, _as_f64_cuda(value))


def tf_one_hot(seq, alphabet, dtype=torch.float64):
    """One-hot encode k-mers -> [n, lag, alphabet_size+1] on the device, start symbol '[' in the last
    column, unknown symbols (protein only) as all-zero rows (core.py:156-174).  ``seq`` is a list /
    array of equal-length strings or bytes, or a ``dataloader.KmerBatch`` (already packed)."""
    from . import dataloader as dl
    if isinstance(seq, dl.KmerBatch):
        packed, lag = seq.packed, seq.lag
    else:
        shape = np.shape(seq)
        codes, lag = dl.encode_kmers(seq, alphabet)
        packed = torch.from_numpy(codes.view(np.int64)).to(_lib.device())
    n = packed.shape[0]
    A1 = dl.ALPHABET_SIZES[alphabet] + 1
    out = torch.empty((n, lag, A1), dtype=torch.float64, device=packed.device)
    check(lib.bear_decode_onehot(ptr(packed), n, lag, _lib.ALPHABET_IDS[alphabet], ptr(out), _lib.stream()))
    if not isinstance(seq, dl.KmerBatch) and len(shape) > 1:
        out = out.reshape(*shape, lag, A1)
    return out.to(dtype)
