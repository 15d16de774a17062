"""world_size-2 CPU (gloo) test of the data-parallel host logic: row sharding of every batch, the
global-batch loss scale, ONE allreduce of the flat [loss, d h, d params] buffer per optimizer step and
identical updates on every rank.  The per-batch gradient is supplied by the oracle here (the CUDA
kernels cannot run on this box); what is under test is bear_b200._engine and the rank-sharded ingest of dataloader."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import YSD1


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _run(rank, world, port, batch, acc_steps, out_path):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    if world > 1:
        dist.init_process_group('gloo', rank=rank, world_size=world)
    from bear_b200 import _engine as eng, dataloader as dl
    from oracle import bear_oracle as O
    torch.set_num_threads(1)
    # under an initialised process group dataloader() parses only this rank's slice of every batch (bear_pack_shard)
    data = dl.dataloader(YSD1, 'dna', batch, 3)
    K = dl.count_rows(YSD1)
    assert sum(g for _, _, g in data.batches()) == K
    calls = {'allreduce': 0}
    if world > 1:
        assert data.table.num_rows < K
        # the same dataset split into two files goes through load_files (what the config scripts call): identical shards
        parts = [os.path.join(os.path.dirname(out_path), 'part_%d.tsv' % i) for i in range(2)]
        if rank == 0:
            lines = open(YSD1).read().splitlines()
            for path, chunk in zip(parts, (lines[:601], lines[601:])):
                with open(path, 'w') as fh:
                    fh.write('\n'.join(chunk) + '\n')
        dist.barrier()
        two = dl.load_files(parts, 'dna', batch, 3)
        assert two.ranges == data.ranges and two.global_rows == data.global_rows and two.row_ids == data.row_ids
        assert np.array_equal(two.table.kmers_host[:two.table.num_rows], data.table.kmers_host[:data.table.num_rows])
        assert np.array_equal(two.table.counts_host[:, :, :two.table.num_rows], data.table.counts_host[:, :, :data.table.num_rows])
        real = dist.all_reduce

        def counting(t, *a, **k):
            calls['allreduce'] += 1
            return real(t, *a, **k)
        dist.all_reduce = counting
    table = data.table
    gen = torch.Generator().manual_seed(0)
    h_signed = torch.zeros((), dtype=torch.float64)
    mat = O.init_linear(5, 4, gen)[0] * 4
    fp = eng.FlatParams([h_signed, mat], device='cpu')
    kmers = [k.decode() for k in table.kmers_str()]
    onehot = O.one_hot(kmers) if kmers else torch.zeros(0, 5, 5, dtype=torch.float64)
    counts = torch.from_numpy(np.transpose(table.counts_host[0, :, :table.num_rows]).astype(np.float64))

    def step_fn(r0, n, scale):
        if n == 0:
            return
        # loss of the local rows with the GLOBAL factor num_kmers / global batch rows
        loss, _, grads = O.train_step_grads(onehot[r0:r0 + n], counts[r0:r0 + n], fp.tensors[0], [fp.tensors[1]],
                                            'linear', scale * n, False)
        fp.grad[0] += loss
        fp.grad[1] += grads[0]
        fp.grad[2:] += grads[1].reshape(-1)

    losses = []
    eng.train_loop(data.repeat(2), K, 0, False, acc_steps, fp, 'SGD', 1e-9, step_fn, loss_save=losses)
    if rank == 0:
        torch.save({'flat': fp.flat.clone(), 'losses': losses, 'allreduce': calls['allreduce'],
                    'steps': len(data.repeat(2))}, out_path)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


@pytest.mark.parametrize('batch,acc_steps', [(500, 1), (300, 2)])
def test_two_rank_training_equals_single_rank(tmp_path, batch, acc_steps):
    single, multi = str(tmp_path / 'single.pt'), str(tmp_path / 'multi.pt')
    _run(0, 1, 0, batch, acc_steps, single)
    mp.spawn(_run, args=(2, _free_port(), batch, acc_steps, multi), nprocs=2, join=True)
    a, b = torch.load(single), torch.load(multi)
    assert len(a['losses']) == len(b['losses']) > 0
    assert np.allclose(a['losses'], b['losses'], rtol=1e-12)
    assert torch.allclose(a['flat'], b['flat'], rtol=1e-12, atol=0)
    # exactly one allreduce per optimizer step
    assert b['allreduce'] == b['steps'] // acc_steps
