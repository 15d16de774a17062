"""ctypes binding of libbear_b200.so (the C-ABI declared in include/bear_b200.h).

There is no fallback: if the shared library has not been built (``python -m bear_b200.build``)
importing this module raises, and every compute entry point requires a CUDA device.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# BEAR_B200_LIB: another build of the same library (kernel experiments, tools/experiments/build_exp.sh)
LIB_PATH = os.environ.get('BEAR_B200_LIB') or os.path.join(_HERE, 'libbear_b200.so')

MAX_MODELS = 8
HEAD_NONE, HEAD_LINEAR, HEAD_EXPLICIT, HEAD_STOP = 0, 1, 2, 3
ALPHABET_IDS = {'dna': 0, 'rna': 1, 'prot': 2}
WIRE_START_ESC = 16               # include/bear_b200.h BEAR_WIRE_START_ESC


class BearError(RuntimeError):
    """A libbear_b200 entry point returned a negative status."""


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            'libbear_b200.so is missing (%s). Build it with `python -m bear_b200.build`; '
            'bear_b200 has no CPU or pure-PyTorch fallback.' % LIB_PATH)
    handle = ctypes.CDLL(LIB_PATH)
    if not os.environ.get('BEAR_B200_LIB'):
        # a binary older than the sources next to it (e.g. after a pull) is refused, never loaded silently;
        # on a box without the sources' compiler the prebuilt library must match the sources it shipped with
        from . import build as _build
        want = _build._digest()
        have = _build.library_digest(LIB_PATH)
        if have != want:
            raise ImportError('libbear_b200.so is stale: built from sources %s, tree is %s. Rebuild with '
                              '`python -m bear_b200.build`.' % (str(have)[:12], want[:12]))
    return handle


lib = _load()

_vp, _i64, _i32, _f64, _cp = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_double, ctypes.c_char_p
_pi64, _pi32 = ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int)

_SIGNATURES = {
    'bear_last_error': (_cp, []),
    'bear_version': (_i32, []),
    'bear_build_digest': (_cp, []),
    'bear_alphabet_size': (_i32, [_i32]),
    'bear_max_lag': (_i32, [_i32]),
    'bear_count_rows': (_i64, [_cp, _i32]),
    'bear_pack_tsv': (_i32, [_cp, _i32, _i32, _i32, _i64, _i64, _vp, _vp, _i64, _pi64, _pi32]),
    'bear_pack_sparse': (_i32, [_cp, _i32, _i32, _i32, _i64, _i64, _vp, _vp, _i64, _pi64, _pi32]),
    'bear_pack_set_invalid_policy': (_i32, [_i32]),
    'bear_pack_shard': (_i32, [_cp, _i32, _i32, _i32, _i32, _i64, _i32, _i32, _i64, _i64, _i64, _vp, _vp, _i64, _pi64, _pi32]),
    'bear_encode_kmers': (_i32, [_cp, _i64, _i32, _i32, _vp]),
    'bear_decode_kmers': (_i32, [_vp, _i64, _i32, _i32, _vp]),
    'bear_compact_bytes': (_i64, [_i64, _i32, _i32, _i32, _i32]),
    'bear_compact_choose_wire': (_i32, [_vp, _vp, _i64, _i64, _i64, _i32, _i32, _i32]),
    'bear_compact_table': (_i32, [_vp, _vp, _i64, _i64, _i64, _i32, _i32, _i32, _i32, _vp, _vp, _i64, _pi64]),
    'bear_expand_table': (_i32, [_vp, _vp, _i64, _i64, _i32, _i32, _i32, _i32, _vp, _vp, _i64, _i64, _vp]),
    'bear_decode_onehot': (_i32, [_vp, _i64, _i32, _i32, _vp, _vp]),
    'bear_decode_symbols': (_i32, [_vp, _i64, _i32, _i32, _vp, _vp]),
    'bear_unpack_counts': (_i32, [_vp, _i64, _i64, _i64, _i32, _i32, _vp, _vp]),
    'bear_dm_logprob': (_i32, [_vp, _i64, _vp, _i64, _i32, _vp, _vp]),
    'bear_dm_logprob_bwd': (_i32, [_vp, _i64, _vp, _i64, _i32, _vp, _vp, _vp]),
    'bear_mn_logprob': (_i32, [_vp, _i64, _vp, _i64, _i32, _vp, _vp]),
    'bear_mn_logprob_bwd': (_i32, [_vp, _i64, _vp, _i64, _i32, _vp, _vp, _vp]),
    'bear_ml_output': (_i32, [_vp, _i64, _i32, _f64, _i64, _vp, _vp]),
    'bear_workspace_doubles': (_i64, [_i64, _i32, _i32]),
    'bear_linear_train_step': (_i32, [_vp, _vp, _i64, _i64, _i64, _i32, _vp, _vp, _f64, _i32, _vp, _vp, _vp, _vp]),
    'bear_dm_train_step_explicit': (_i32, [_vp, _i64, _i64, _i64, _vp, _vp, _f64, _i32, _vp, _vp, _vp, _vp, _vp]),
    'bear_eval_step': (_i32, [_vp, _vp, _vp, _i64, _i64, _i64, _i32, _i32, _vp, _vp, _i32, _vp, _i32, _i64, _i64, _vp, _vp, _vp]),
    'bear_bmm_likelihood': (_i32, [_vp, _i64, _i64, _i64, _i32, _i32, _vp, _i32, _vp, _vp, _vp]),
    'bear_loggamma_sample': (_i32, [_vp, _i64, _i64, _i64, _vp, _vp]),
    'bear_log_normalize': (_i32, [_vp, _i64, _i32, _vp]),
    'bear_synth_table': (_i32, [_vp, _vp, _i64, _i64, _i64, _i32, _i32, _i64, _i32, _i32, _vp]),
    'bear_count_transitions': (_i32, [_vp, _vp, _vp, _vp, _i64, _i64, _i32, _i32, _i32, _vp, _vp, _i64, _vp, _vp]),
    'bear_gather_table': (_i32, [_vp, _vp, _vp, _i64, _i32, _i64, _vp, _vp, _vp]),
    'bear_adam_step': (_i32, [_vp, _vp, _vp, _vp, _i64, _f64, _f64, _f64, _f64, _vp, _vp, _f64, _i32, _vp]),
    'bear_adam_update': (_i32, [_vp, _vp, _vp, _vp, _i64, _f64, _f64, _f64, _f64, _vp, _vp]),
    'bear_ref_train_step': (_i32, [_vp, _vp, _i64, _i64, _i64, _vp, _vp, _vp, _f64, _i32, _vp, _vp, _vp, _vp]),
    'bear_ref_eval_step': (_i32, [_vp, _vp, _vp, _vp, _i64, _i64, _i64, _i32, _i32, _vp, _vp, _vp, _vp, _i32, _vp, _i32, _i64, _i64, _vp,
                                  _vp, _vp]),
    'bear_cnn_supported': (_i32, [_i32, _i32, _i32, _i32]),
    'bear_cnn_num_params': (_i64, [_i32, _i32, _i32, _i32]),
    'bear_cnn_head_forward': (_i32, [_vp, _i64, _i64, _i32, _i32, _i32, _i32, _vp, _vp, _vp]),
    'bear_cnn_train_step': (_i32, [_vp, _vp, _i64, _i64, _i64, _i32, _i32, _i32, _i32, _vp, _vp, _f64, _i32, _vp, _vp,
                                   _vp, _vp]),
    'bear_cnn_head_backward': (_i32, [_vp, _i64, _i64, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp]),
    'bear_ref_head': (_i32, [_vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp]),
    'bear_ref_head_bwd': (_i32, [_vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
}

EXPORTS = sorted(_SIGNATURES)

for _name, (_res, _args) in _SIGNATURES.items():
    _fn = getattr(lib, _name)     # AttributeError here = header / library mismatch
    _fn.restype = _res
    _fn.argtypes = _args


def last_error():
    return lib.bear_last_error().decode('utf-8', 'replace')


def check(rc):
    if rc < 0:
        raise BearError('libbear_b200: %s (status %d)' % (last_error(), rc))
    return rc


def ptr(t):
    """Raw address of a tensor / numpy array (None -> NULL)."""
    if t is None:
        return None
    if isinstance(t, torch.Tensor):
        return ctypes.c_void_p(t.data_ptr())
    return ctypes.c_void_p(t.ctypes.data)


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise BearError('bear_b200 kernels need CUDA tensors; got a %s tensor (no CPU fallback)' % t.device)


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def device():
    """The CUDA device this process computes on (one process per GPU)."""
    if not torch.cuda.is_available():
        raise BearError('no CUDA device is available; bear_b200 has no CPU fallback')
    return torch.device('cuda', torch.cuda.current_device())
