"""Two NCCL ranks on two GPUs against the single-GPU run, through the public API (skipped when fewer than two GPUs are
visible): bear_net.train -- fused tcgen05 train kernel, ONE allreduce of the flat buffer per optimizer step, one optimizer
launch, also when the epoch is replayed from a CUDA graph -- and bear_net.evaluation, whose tie-break noise is keyed on
the global row so that even the integer accuracy counts agree exactly."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _run(rank, world, port, epochs, batch, out_path):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    if world > 1:
        dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    from bear_b200 import ar_funcs, bear_net, dataloader as dl
    lag, K = 13, 60000
    rng = np.random.default_rng(21)
    codes = rng.integers(0, 4 ** lag, size=K, dtype=np.uint64)
    ns = np.where(rng.random(K) < 0.05, rng.integers(1, lag + 1, size=K), 0).astype(np.uint64)
    codes = (codes & ((np.uint64(1) << (np.uint64(2) * (np.uint64(lag) - ns))) - np.uint64(1))) | (ns << np.uint64(58))
    counts = rng.poisson(0.6, size=(K, 2, 5)).astype(np.int64) * (rng.random((K, 2, 1)) < 0.8)
    data = dl.KmerDataset(dl.KmerTable.from_arrays((codes, lag), counts, 'dna'), batch)
    if world > 1:
        data = data.shard(rank, world)
    torch.manual_seed(5)
    p0, _, _ = bear_net._create_params(lag, 4, ar_funcs.make_ar_func_linear, {})
    p0 = [p.clone() for p in p0]
    losses = []
    params, h_signed, ar_func = bear_net.train(data.repeat(epochs), K, epochs, 0, 'dna', lag, ar_funcs.make_ar_func_linear, {},
                                               0.01, 'Adam', False, params_restart=p0, loss_save=losses)
    ev = bear_net.evaluation(data, 0, 1, 'dna', torch.exp(h_signed), ar_func, np.array([0.1, 1.0, 10.0]), seed=77)
    torch.cuda.synchronize()
    if rank == 0:
        torch.save({'params': [p.cpu() for p in params], 'losses': losses, 'eval': [torch.as_tensor(e).cpu() for e in ev]}, out_path)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


@pytest.mark.parametrize('epochs,batch,graph', [(3, 7000, False), (40, 20000, True)])
def test_two_nccl_ranks_equal_one_rank(tmp_path, monkeypatch, epochs, batch, graph):
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs')
    # graph = True: low launch threshold, so that the epochs after the first replay from a CUDA graph (kernels + NCCL)
    monkeypatch.setenv('BEAR_GRAPH_MIN_LAUNCHES', '16' if graph else '1000000000')
    single, multi = str(tmp_path / 'single.pt'), str(tmp_path / 'multi.pt')
    mp.spawn(_run, args=(1, _free_port(), epochs, batch, single), nprocs=1, join=True)
    mp.spawn(_run, args=(2, _free_port(), epochs, batch, multi), nprocs=2, join=True)
    a, b = torch.load(single), torch.load(multi)
    assert len(a['losses']) == len(b['losses']) == epochs * -(-60000 // batch)
    assert np.allclose(a['losses'], b['losses'], rtol=1e-11)
    for pa, pb in zip(a['params'], b['params']):
        assert torch.allclose(pa, pb, rtol=1e-9, atol=1e-12)
    for i, (ea, eb) in enumerate(zip(a['eval'], b['eval'])):
        assert torch.allclose(ea, eb, rtol=1e-10, atol=0), (i, ea, eb)
    # accuracies are ratios of integer counts: the same rows pick the same letters on one and on two ranks
    for i in (6, 7, 8):
        assert torch.equal(a['eval'][i], b['eval'][i])
