"""Pluggable autoregressive heads.

Mirrors the reference's ``bear_model/ar_funcs.py`` plugin convention: ``make_ar_func_<name>(lag,
alphabet_size, **af_kwargs, dtype) -> (ar_func, params)``, looked up by
``getattr(ar_funcs, 'make_ar_func_' + name)`` (models/train_bear_net.py:103).  ``ar_func`` maps a
one-hot tensor ``[..., lag, alphabet_size+1]`` to transition probabilities ``[..., alphabet_size+1]``.

The returned callables additionally carry ``kind`` and ``params`` so that ``bear_net`` / ``bear_ref``
can route the built-in heads to the fused packed-path kernels (the linear head never materialises a
one-hot tensor there).  Any other callable -- a user plugin written with torch ops -- is served by
the explicit-head path: device decode -> plugin -> ``bear_dm_train_step_explicit``.
"""
import torch

from . import _lib


def _normalize_layer(layer, reduce_dims=(-1,)):
    """(x - mean) / sqrt(biased var + 1e-5) over ``reduce_dims`` (ar_funcs.py:5-20)."""
    dims = tuple(reduce_dims)
    mean = layer.mean(dim=dims, keepdim=True)
    var = ((layer - mean) ** 2).mean(dim=dims, keepdim=True)
    return (layer - mean) / torch.sqrt(var + 1E-5)


def _l2_normalize(x, dims):
    return x / torch.sqrt((x * x).sum(dim=dims, keepdim=True).clamp_min(1e-12))


class ARFunc:
    """Callable head with its parameter list.  ``kind`` in {'linear', 'cnn', 'stop', 'custom'}."""
    kind = 'custom'

    def __init__(self, params):
        self.params = list(params)


class LinearARFunc(ARFunc):
    """softmax(einsum('...jk,jkl->...l', kmers, mat)) (ar_funcs.py:44-45)."""
    kind = 'linear'

    def __call__(self, kmers):
        return torch.softmax(torch.einsum('...jk,jkl->...l', kmers, self.params[0]), dim=-1)


class CNNARFunc(ARFunc):
    """conv1d(VALID) -> layer-norm -> elu -> dense -> layer-norm -> elu -> dense -> softmax
    (ar_funcs.py:91-97); params in the reference order (ar_funcs.py:98-99)."""
    kind = 'cnn'

    def __call__(self, data):
        filters, int0, W1, int1, W2, int2, scale0, scale1 = self.params
        W = filters.shape[0]
        win = data.unfold(-2, W, 1)                                  # [..., P, A1, W]
        conv = torch.einsum('...paw,waf->...pf', win, filters)
        x0 = scale0 * _normalize_layer(conv) + int0
        x1 = scale1 * _normalize_layer(torch.einsum('...pf,pfh->...h', torch.nn.functional.elu(x0), W1)) + int1
        x2 = torch.nn.functional.elu(x1) @ W2 + int2
        return torch.softmax(x2, dim=-1)


class StopARFunc(ARFunc):
    """Always predicts a stop: the constant [0, ..., 0, 1] (ar_funcs.py:121-126)."""
    kind = 'stop'

    def __init__(self, stop):
        super().__init__([])
        self.stop = stop

    def __call__(self, y):
        return self.stop


def make_ar_func_linear(lag, alphabet_size, dtype=torch.float64):
    """Linear autoregressive function (ar_funcs.py:23-46).  params = [mat [lag, A+1, A+1]],
    initialised to 0.05 * l2_normalize(N(0,1), axis=1)."""
    dev = _lib.device()
    mat = torch.randn(lag, alphabet_size + 1, alphabet_size + 1, dtype=dtype, device=dev)
    mat = 0.05 * _l2_normalize(mat, (1,))
    return LinearARFunc([mat]), [mat]


def make_ar_func_cnn(lag, alphabet_size, filter_width=8, num_filters=30, kmer_layer1_width=16,
                     dtype=torch.float64):
    """Convolutional autoregressive function (ar_funcs.py:49-99).  params =
    [filters, intercept0, weights1, intercept1, weights2, intercept2, scale0, scale1]."""
    dev = _lib.device()
    W, F, H1 = int(filter_width), int(num_filters), int(kmer_layer1_width)
    a1 = alphabet_size + 1
    P = lag - W + 1
    if P < 1:
        raise ValueError('filter_width %d exceeds lag %d' % (W, lag))
    kw = dict(dtype=dtype, device=dev)
    filters = _l2_normalize(torch.randn(W, a1, F, **kw), (0, 1))
    int0 = torch.ones(P, F, **kw)
    scale0 = torch.ones(P, F, **kw)
    W1 = _l2_normalize(torch.randn(P, F, H1, **kw), (0,))
    int1 = torch.ones(H1, **kw)
    scale1 = torch.ones(H1, **kw)
    W2 = 0.05 * _l2_normalize(torch.randn(H1, a1, **kw), (0,))
    int2 = torch.zeros(a1, **kw)
    params = [filters, int0, W1, int1, W2, int2, scale0, scale1]
    return CNNARFunc(params), params


def make_ar_func_stop(lag, alphabet_size, dtype=torch.float64):
    """Head that always predicts a stop; for use with the reference AR model (ar_funcs.py:102-127)."""
    stop = torch.zeros(alphabet_size + 1, dtype=dtype, device=_lib.device())
    stop[-1] = 1
    return StopARFunc(stop), []
