"""Sampling log X for X ~ Gamma(conc, 1), stable for tiny concentrations.

Mirrors the reference's ``bear_model/log_gamma.py`` (``log_gamma`` log_gamma.py:17-76,
``log_gamma_pdf`` :14-15, ``triple_slice_mut`` :5-12).  The reference rejection-samples on the host
with NumPy; here every draw is produced by ``bear_loggamma_sample`` on the device with a
counter-based generator: Marsaglia-Tsang for conc >= 1 and log Gamma(conc + 1) + log(U)/conc below,
which is the same distribution (the reference's own test is a KS test, tests/test_log_gamma.py:9-19).
"""
import numpy as np
import torch

from . import _lib
from ._lib import lib, check, ptr


def triple_slice_mut(b, c, d):
    """Boolean mask equal to the slice a[b][c][d] (log_gamma.py:5-12)."""
    change = np.zeros(len(b), dtype=bool)
    change_temp = np.zeros(len(c), dtype=bool)
    change_temp[c] = d
    change[b] = change_temp
    return change


def log_gamma_pdf(conc, xs):
    """Density of log X, X ~ Gamma(conc, 1) (log_gamma.py:14-15)."""
    xs = np.asarray(xs, dtype=np.float64)
    lg = torch.lgamma(torch.as_tensor(np.asarray(conc, dtype=np.float64))).numpy()
    return np.exp(conc * xs - np.exp(xs) - lg)


def log_gamma_device(concs, n_samples, seed=None):
    """Device tensor in, device tensor out: [n_samples, *concs.shape]."""
    concs = concs.to(_lib.device(), torch.float64).contiguous()
    if seed is None:
        seed = int(torch.randint(0, 2 ** 62, (1,)).item())
    out = torch.empty((n_samples, concs.numel()), dtype=torch.float64, device=concs.device)
    check(lib.bear_loggamma_sample(ptr(concs), concs.numel(), n_samples, seed, ptr(out), _lib.stream()))
    return out.reshape(n_samples, *concs.shape)


def log_gamma(concs, size=[], seed=None):
    """Samples of shape ``size + shape(concs)`` (numpy), log_gamma.py:17-76."""
    concs = np.asarray(concs, dtype=np.float64)
    size = [int(s) for s in np.atleast_1d(size)] if np.size(size) else []
    n_samples = int(np.prod(size)) if size else 1
    out = log_gamma_device(torch.from_numpy(concs), n_samples, seed)
    return out.cpu().numpy().reshape(size + list(concs.shape))
