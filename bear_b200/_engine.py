"""Shared machinery of bear_net / bear_ref: flat parameter + gradient buffers, the per-step
allreduce, Keras-style optimizers, and the routing of AR heads onto the fused kernels.

Data parallelism (replaces tf.distribute.MirroredStrategy, bear_net.py:246,273,290-291): one process
per GPU, k-mer rows sharded by ``KmerDataset.shard``; per optimizer step ONE allreduce(SUM) of the
flat float64 buffer ``[loss, d h_signed, d params...]`` (<= 52 KB), then the identical update on
every rank.  Evaluation accumulates locally and allreduces its handful of scalars once at the end.
"""
import math
import os

import numpy as np
import torch
import torch.distributed as dist

from . import _lib, core
from ._lib import lib, check, ptr
from .ar_funcs import ARFunc
from .dataloader import KmerDataset

EXPLICIT_CHUNK = 1 << 18      # rows per one-hot materialisation for explicit (torch-op) heads


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def allreduce_sum(t):
    if world()[1] > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def head_kind(ar_func):
    return getattr(ar_func, 'kind', 'custom') if ar_func is not None else 'none'


def fused_linear_ok(ar_func, table):
    return head_kind(ar_func) == 'linear' and table.alphabet in ('dna', 'rna')


def cnn_dims(ar_func):
    """(filter_width, num_filters, kmer_layer1_width) of a CNN head from its parameter shapes (ar_funcs.py:78-89)."""
    filters, W1 = ar_func.params[0], ar_func.params[2]
    return int(filters.shape[0]), int(filters.shape[2]), int(W1.shape[2])


def fused_cnn_ok(ar_func, table):
    """True when the CNN head runs in the hand-written fused kernels (bear_cnn_*): DNA / RNA table and
    dimensions inside the kernels' envelope (F <= 32, H1 <= 16, tile fits shared memory)."""
    if head_kind(ar_func) != 'cnn' or table.alphabet not in ('dna', 'rna') or os.environ.get('BEAR_CNN_TORCH'):
        return False
    W, F, H1 = cnn_dims(ar_func)
    return bool(lib.bear_cnn_supported(table.lag, W, F, H1))


def cnn_param_block(params):
    """The eight CNN parameter arrays as one contiguous float64 block in the reference's list order
    (ar_funcs.py:98-99).  Views of a FlatParams buffer already are one; anything else is concatenated."""
    consecutive = all(p.is_contiguous() and p.dtype == torch.float64 for p in params) and all(
        params[i + 1].data_ptr() == params[i].data_ptr() + 8 * params[i].numel() for i in range(len(params) - 1))
    if consecutive:
        n = sum(int(p.numel()) for p in params)
        return torch.as_strided(params[0].detach(), (n,), (1,))
    return torch.cat([p.detach().reshape(-1).to(torch.float64) for p in params]).contiguous()


def cnn_forward(ar_func, table, r0, n, block=None):
    """f[n, 5] = CNN head of rows [r0, r0+n) straight from the packed codes (bear_cnn_head_forward)."""
    k, _ = table.device_tensors()
    W, F, H1 = cnn_dims(ar_func)
    if block is None:
        block = cnn_param_block(ar_func.params)
    f = torch.empty((n, table.A1), dtype=torch.float64, device=k.device)
    check(lib.bear_cnn_head_forward(ptr(k), r0, n, table.lag, W, F, H1, ptr(block), ptr(f), _lib.stream()))
    return f


class _CnnPacked(torch.autograd.Function):
    """Differentiable f = CNN(packed rows) for callers that put more torch / kernel stages after the head
    (bear_ref embeds it as the net g, bear_ref.py:63-68): forward and backward are the fused kernels."""

    @staticmethod
    def forward(ctx, block, ar_func, table, r0, n):
        ctx.args = (ar_func, table, r0, n)
        ctx.save_for_backward(block)
        return cnn_forward(ar_func, table, r0, n, block.detach().contiguous())

    @staticmethod
    def backward(ctx, gf):
        ar_func, table, r0, n = ctx.args
        block, = ctx.saved_tensors
        k, _ = table.device_tensors()
        W, F, H1 = cnn_dims(ar_func)
        gblock = torch.zeros_like(block)
        ws = workspace(table, block.numel())
        check(lib.bear_cnn_head_backward(ptr(k), r0, n, table.lag, W, F, H1, ptr(block.detach().contiguous()),
                                         ptr(gf.contiguous()), ptr(gblock), ptr(ws), _lib.stream()))
        return gblock, None, None, None, None


def cnn_forward_autograd(ar_func, table, r0, n):
    block = torch.cat([p.reshape(-1) for p in ar_func.params])
    return _CnnPacked.apply(block, ar_func, table, r0, n)


class FlatParams:
    """[h_signed, params...] packed in one contiguous float64 device buffer.  The tensors handed to
    the user (and captured by plugin closures) are re-pointed to views of it, so the optimizer
    kernel and the allreduce see a single buffer while ``ar_func`` keeps working unchanged."""

    def __init__(self, tensors, device=None):
        dev = _lib.device() if device is None else torch.device(device)
        self.tensors = list(tensors)
        self.sizes = [int(t.numel()) for t in self.tensors]
        self.total = sum(self.sizes)
        self.flat = torch.zeros(self.total, dtype=torch.float64, device=dev)
        self.offsets = []
        o = 0
        for t, n in zip(self.tensors, self.sizes):
            view = self.flat[o:o + n].view(t.shape)
            view.copy_(t.detach().to(dev, torch.float64))
            t.data = view
            self.offsets.append(o)
            o += n
        # gradient buffer: [loss, d params...]
        self.grad = torch.zeros(1 + self.total, dtype=torch.float64, device=dev)

    def grad_view(self, i):
        o = 1 + self.offsets[i]
        return self.grad[o:o + self.sizes[i]].view(self.tensors[i].shape)


class Optimizer:
    """tf.keras.optimizers.<name>(learning_rate) with Keras (OptimizerV2) defaults, on the flat buffer
    (bear_net.py:264-265: ``getattr(tf.keras.optimizers, optimizer_name)``).  Adam runs as one libbear_b200 kernel; SGD,
    RMSprop, Adagrad, Adadelta, Adamax and Nadam are a few torch ops on the flat buffer (Ftrl is not implemented)."""

    def __init__(self, name, learning_rate, n, device):
        self.name, self.lr = name, float(learning_rate)
        z = lambda: torch.zeros(n, dtype=torch.float64, device=device)
        if name == 'Adam':
            self.m, self.v = z(), z()
            self.step = torch.zeros(1, dtype=torch.int64, device=device)
        elif name == 'SGD':
            pass
        elif name == 'RMSprop':
            self.v = z()
        elif name == 'Adagrad':
            self.v = z() + 0.1
        elif name == 'Adadelta':
            self.v, self.d = z(), z()                      # accumulated squared gradients / squared updates
        elif name == 'Adamax':
            self.m, self.v = z(), z()                      # first moment, exponentially weighted infinity norm
            self.t = torch.zeros((), dtype=torch.float64, device=device)
        elif name == 'Nadam':
            self.m, self.v = z(), z()
            self.t = torch.zeros((), dtype=torch.float64, device=device)
            self.msched = torch.ones((), dtype=torch.float64, device=device)     # product of the momentum schedule
        else:
            raise ValueError("optimizer '%s' is not supported (Adam, SGD, RMSprop, Adagrad, Adadelta, Adamax, Nadam)" % name)

    def apply(self, params, grads):
        if self.name == 'Adam':
            check(lib.bear_adam_update(ptr(params), ptr(grads), ptr(self.m), ptr(self.v), params.numel(),
                                       self.lr, 0.9, 0.999, 1e-7, ptr(self.step), _lib.stream()))
        elif self.name == 'SGD':
            params.sub_(self.lr * grads)
        elif self.name == 'RMSprop':
            self.v.mul_(0.9).addcmul_(grads, grads, value=0.1)
            params.sub_(self.lr * grads / (self.v.sqrt() + 1e-7))
        elif self.name == 'Adagrad':
            self.v.addcmul_(grads, grads)
            params.sub_(self.lr * grads / (self.v.sqrt() + 1e-7))
        elif self.name == 'Adadelta':                      # rho = 0.95, epsilon = 1e-7 (Keras defaults)
            self.v.mul_(0.95).addcmul_(grads, grads, value=0.05)
            upd = grads * ((self.d + 1e-7).sqrt() / (self.v + 1e-7).sqrt())
            self.d.mul_(0.95).addcmul_(upd, upd, value=0.05)
            params.sub_(self.lr * upd)
        elif self.name == 'Adamax':                        # beta_1 = 0.9, beta_2 = 0.999, epsilon = 1e-7
            self.t += 1.0                                  # (step counters are device scalars: graph-capturable)
            self.m.mul_(0.9).add_(grads, alpha=0.1)
            torch.maximum(self.v * 0.999, grads.abs(), out=self.v)
            params.sub_((self.lr / (1.0 - torch.pow(0.9, self.t))) * self.m / (self.v + 1e-7))
        elif self.name == 'Nadam':                         # Keras: momentum schedule u_t = beta_1 (1 - 0.5 * 0.96^(0.004 t))
            self.t += 1.0
            u_t = 0.9 * (1.0 - 0.5 * torch.pow(0.96, 0.004 * self.t))
            u_n = 0.9 * (1.0 - 0.5 * torch.pow(0.96, 0.004 * (self.t + 1.0)))
            self.msched.mul_(u_t)
            self.m.mul_(0.9).add_(grads, alpha=0.1)
            self.v.mul_(0.999).addcmul_(grads, grads, value=0.001)
            g_hat = grads / (1.0 - self.msched)
            m_hat = self.m / (1.0 - self.msched * u_n)
            v_hat = self.v / (1.0 - torch.pow(0.999, self.t))
            params.sub_(self.lr * ((1.0 - u_t) * g_hat + u_n * m_hat) / (v_hat.sqrt() + 1e-7))

    def step_and_clear(self, fp, loss_slot, loss_scale):
        """One optimizer step on the (already allreduced) flat buffer ``fp.grad`` = [loss, grads...]: records
        ``loss_scale * loss`` into ``loss_slot`` (a one-element device view, or None), updates ``fp.flat`` and
        clears ``fp.grad`` for the next accumulation.  Adam: one kernel launch (bear_adam_step)."""
        if self.name == 'Adam':
            check(lib.bear_adam_step(ptr(fp.flat), ptr(fp.grad), ptr(self.m), ptr(self.v), fp.total, self.lr, 0.9, 0.999,
                                     1e-7, ptr(self.step), ptr(loss_slot) if loss_slot is not None else None,
                                     float(loss_scale), 1, _lib.stream()))
            return
        if loss_slot is not None:
            loss_slot.copy_(loss_scale * fp.grad[:1])
        self.apply(fp.flat, fp.grad[1:])
        fp.grad.zero_()


def workspace(table, nparams=0):
    return torch.empty(lib.bear_workspace_doubles(table.num_rows, table.lag, nparams), dtype=torch.float64,
                       device=_lib.device())


def onehot_rows(table, r0, n):
    """core.tf_one_hot of rows [r0, r0+n) from the packed codes, on the device."""
    k, _ = table.device_tensors()
    out = torch.empty((n, table.lag, table.A1), dtype=torch.float64, device=k.device)
    check(lib.bear_decode_onehot(ptr(k[r0:r0 + n]), n, table.lag, _lib.ALPHABET_IDS[table.alphabet], ptr(out),
                                 _lib.stream()))
    return out


def dense_counts(table, r0, n):
    """The reference-shaped counts tensor [n, G, A1] (float64) of rows [r0, r0+n), on the device."""
    k, c = table.device_tensors()
    out = torch.empty((n, table.num_ds, table.A1), dtype=torch.float64, device=k.device)
    check(lib.bear_unpack_counts(ptr(c), table.stride, r0, n, table.num_ds, table.A1, ptr(out), _lib.stream()))
    return out


def check_dataset(data):
    if not isinstance(data, KmerDataset):
        raise TypeError('data must be a KmerDataset from bear_b200.dataloader (dataloader / sparse_dataloader)')
    if data.map_fn is not None:
        raise TypeError('train / evaluation take the un-mapped KmerDataset; column selection happens on the device')
    return data.table


# ---------------------------------------------------------------------------------------------
# training
# ---------------------------------------------------------------------------------------------
def train_loop(data, num_kmers, ds_loc, train_ar, acc_steps, fp, optimizer_name, learning_rate, step_fn,
               writer=None, loss_save=None, graph_safe=False):
    """The loop of bear_net.train (bear_net.py:293-315): ``step_fn(r0, n, scale)`` adds the loss and
    gradients of one batch into ``fp.grad``; every ``acc_steps`` batches the buffer is allreduced,
    the loss recorded and the optimizer applied (one launch)."""
    opt = Optimizer(optimizer_name, learning_rate, fp.total, fp.flat.device)
    fp.grad.zero_()
    record = writer is not None or loss_save is not None
    nb, reps = len(data.ranges), data.repeats
    upd = nb // acc_steps                 # optimizer steps per epoch when nb is a multiple of acc_steps
    dev = fp.flat.device

    def one_epoch(loss_buf, first_step=0, slot0=0):
        """Every batch of one pass over the data; per optimizer step (every ``acc_steps`` batches, counted over the whole
        run as in bear_net.py:300): the batch kernels, ONE allreduce of the flat buffer, ONE optimizer launch (loss
        record + update + clearing of the buffer).  Returns the number of optimizer steps taken."""
        u = 0
        for i, ((r0, n), gB) in enumerate(zip(data.ranges, data.global_rows)):
            step_fn(r0, n, float(num_kmers) / float(gB))
            if (first_step + i + 1) % acc_steps == 0:
                allreduce_sum(fp.grad)
                opt.step_and_clear(fp, loss_buf[slot0 + u:slot0 + u + 1] if loss_buf is not None else None, -1.0 / acc_steps)
                u += 1
        return u

    # Small batches make this loop launch-bound (the reference's tutorial: 10 000 steps over 1 365 rows; a 2^22-row
    # global batch on 8 GPUs is 27 us of kernel time per step), so when the step only launches libbear_b200 kernels
    # (``graph_safe``) one epoch -- batch kernels, NCCL allreduce and optimizer launch of every step -- is captured in a
    # CUDA graph after an eager first epoch and replayed for the remaining ones.  Capture costs ~0.2 s, so only runs
    # that replay >= BEAR_GRAPH_MIN_LAUNCHES launches use it.
    min_launches = int(os.environ.get('BEAR_GRAPH_MIN_LAUNCHES', 8192))
    use_graph = (graph_safe and fp.flat.is_cuda and reps >= 2 and 0 < nb <= 4096 and nb % acc_steps == 0
                 and (reps - 1) * nb * 4 >= min_launches and not os.environ.get('BEAR_NO_GRAPH'))
    if world()[1] > 1 and os.environ.get('BEAR_NO_NCCL_GRAPH'):
        use_graph = False
    losses = []
    if use_graph:
        loss_buf = torch.zeros(max(upd, 1), dtype=torch.float64, device=dev)
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):           # eager first epoch (also warms up everything capture needs)
            one_epoch(loss_buf)
        cur.wait_stream(side)
        if record:
            losses.append(loss_buf.clone())
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            one_epoch(loss_buf)
        for _ in range(1, reps):
            graph.replay()
            if record:
                losses.append(loss_buf.clone())
        update_steps = [(e * nb) + (u + 1) * acc_steps for e in range(reps) for u in range(upd)]
    else:
        total_upd = (reps * nb) // acc_steps
        all_losses = torch.zeros(max(total_upd, 1), dtype=torch.float64, device=dev) if record else None
        done = 0
        for e in range(reps):
            done += one_epoch(all_losses, e * nb, done)
        losses = [all_losses[:done]] if record else []
        update_steps = [(u + 1) * acc_steps for u in range(done)]
    if record and update_steps:
        for s_, v_ in zip(update_steps, torch.cat(losses).cpu().tolist()):     # one D2H copy at the end
            if loss_save is not None:
                loss_save.append(v_)
            if writer is not None and hasattr(writer, 'add_scalar'):
                writer.add_scalar('elbo', v_, s_)


def explicit_f(ar_func, table, r0, n, extra=None):
    """f = ar_func(onehot) (and for bear_ref heads, of the reference counts) for rows [r0, r0+n)."""
    oh = onehot_rows(table, r0, n)
    f = ar_func(oh) if extra is None else ar_func(oh, extra)
    if f.dim() == 1:
        f = f.expand(n, -1)
    return f


def explicit_train_step(table, ds_loc, r0, n, scale, train_ar, fp, ws, f_fn):
    """One batch through a caller-evaluated head: chunks of rows -> f (torch autograd) ->
    bear_dm_train_step_explicit -> backward through the head."""
    h_signed = fp.tensors[0]
    for c0 in range(r0, r0 + n, EXPLICIT_CHUNK):
        cn = min(EXPLICIT_CHUNK, r0 + n - c0)
        with torch.enable_grad():
            f = f_fn(c0, cn)
        fc = f.detach().contiguous()
        gf = torch.empty_like(fc)
        check(lib.bear_dm_train_step_explicit(table.col_ptr(ds_loc), table.stride, c0, cn, ptr(fc), ptr(h_signed),
                                              scale, int(train_ar), ptr(fp.grad), ptr(gf), None, ptr(ws),
                                              _lib.stream()))
        if f.requires_grad:
            f.backward(gf)
    for i, t in enumerate(fp.tensors):
        if t.grad is not None:
            fp.grad_view(i).add_(t.grad)
            t.grad = None


# ---------------------------------------------------------------------------------------------
# evaluation
# ---------------------------------------------------------------------------------------------
def eval_loop(data, ds_loc_train, ds_loc_test, h, van_reg, head, head_ptr_fn, seed, ref=None):
    """bear_net.evaluation's accumulation (bear_net.py:439-457) on the packed table.  ``head`` is a
    BEAR_HEAD_* id; ``head_ptr_fn(r0, n)`` returns (tensor kept alive, explicit row offset) for the
    head argument of bear_eval_step.  ``ref`` = (ds_loc_ref, net id, mat or None, tau_signed, net_weight_signed) routes
    the pass through the fused reference-head kernel bear_ref_eval_step instead.  Returns float64 CPU tensors
    (ll_ear[H], ll_arm, ll_van[V], cor_ear[H], cor_arm, cor_van[V], total_len)."""
    table = data.table
    k, c = table.device_tensors()
    dev = k.device
    h = torch.as_tensor(np.asarray(h, dtype=np.float64)).reshape(-1)
    van = torch.as_tensor(np.asarray(van_reg, dtype=np.float64)).reshape(-1)
    H, V = h.numel(), van.numel()
    if seed is None:
        seed = int(torch.randint(0, 2 ** 62, (1,)).item())
    ws = workspace(table)
    test_ptr = table.col_ptr(ds_loc_test)
    train_ptr = table.col_ptr(ds_loc_train) if ds_loc_train >= 0 else None
    M = _lib.MAX_MODELS
    passes = max(1, -(-H // M), -(-V // M))
    ll_ear, cor_ear = torch.zeros(H, dtype=torch.float64), torch.zeros(H, dtype=torch.float64)
    ll_van, cor_van = torch.zeros(V, dtype=torch.float64), torch.zeros(V, dtype=torch.float64)
    ll_arm = cor_arm = total = None
    for p in range(passes):
        hp = h[p * M:(p + 1) * M].to(dev)
        vp = van[p * M:(p + 1) * M].to(dev)
        Hp, Vp = hp.numel(), vp.numel()
        acc = torch.zeros(2 * Hp + 2 * Vp + 3, dtype=torch.float64, device=dev)
        for (r0, n), gid in list(zip(data.ranges, data.row_ids)) * data.repeats:
            if ref is not None:
                ds_loc_ref, net, mat, tau_signed, nw_signed = ref
                check(lib.bear_ref_eval_step(ptr(k), test_ptr, train_ptr, table.col_ptr(ds_loc_ref), table.stride, r0, n,
                                             table.lag, net, ptr(mat), ptr(tau_signed), ptr(nw_signed),
                                             ptr(hp) if Hp else None, Hp, ptr(vp) if Vp else None, Vp, seed, gid, ptr(acc),
                                             ptr(ws), _lib.stream()))
                continue
            keep, hptr = head_ptr_fn(r0, n)
            check(lib.bear_eval_step(ptr(k), test_ptr, train_ptr, table.stride, r0, n, table.lag, head, hptr,
                                     ptr(hp) if Hp else None, Hp, ptr(vp) if Vp else None, Vp, seed, gid, ptr(acc),
                                     ptr(ws), _lib.stream()))
            del keep
        acc = allreduce_sum(acc).cpu()
        o = 0
        ll_ear[p * M:p * M + Hp] = acc[o:o + Hp]; o += Hp
        if p == 0:
            ll_arm = acc[o].clone()
        o += 1
        ll_van[p * M:p * M + Vp] = acc[o:o + Vp]; o += Vp
        cor_ear[p * M:p * M + Hp] = acc[o:o + Hp]; o += Hp
        if p == 0:
            cor_arm = acc[o].clone()
        o += 1
        cor_van[p * M:p * M + Vp] = acc[o:o + Vp]; o += Vp
        if p == 0:
            total = acc[o].clone()
    return ll_ear, ll_arm, ll_van, cor_ear, cor_arm, cor_van, total


def explicit_head_ptr_fn(f_fn):
    """Adapter for heads evaluated with torch ops: computes f for the batch and passes it as a
    BEAR_HEAD_EXPLICIT array whose row 0 is table row r0."""
    def fn(r0, n):
        with torch.no_grad():
            parts = [f_fn(c0, min(EXPLICIT_CHUNK, r0 + n - c0)) for c0 in range(r0, r0 + n, EXPLICIT_CHUNK)]
            f = (parts[0] if len(parts) == 1 else torch.cat(parts)).contiguous()
        return f, ptr(f)
    return fn


def finish_evaluation(ll_ear, ll_arm, ll_van, cor_ear, cor_arm, cor_van, total):
    """bear_net.py:459-463."""
    return (ll_ear, ll_arm, ll_van,
            torch.exp(-ll_ear / total), torch.exp(-ll_arm / total), torch.exp(-ll_van / total),
            cor_ear / total, cor_arm / total, cor_van / total)
