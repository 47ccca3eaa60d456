"""Data-parallel host logic on CPU: world_size-2 ``gloo`` run of istgcn.dp.GradBuckets (bucket
cutting in reverse order, overlap hooks, deferred mode used under CUDA graphs, exclusion of
never-used parameters, averaging) against the single-process mean of the per-rank gradients."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn


class Tiny(nn.Module):
    def __init__(self):
        super().__init__()
        self.a = nn.Linear(6, 16)
        self.gcn = nn.Module()
        self.gcn.branch = nn.Module()
        self.gcn.branch.bn = nn.BatchNorm1d(4)          # registered, never used (Inception2.bn)
        self.b = nn.Linear(16, 8)
        self.c = nn.Linear(8, 3)

    def forward(self, x):
        return self.c(torch.relu(self.b(torch.relu(self.a(x)))))


def _data(rank):
    g = torch.Generator().manual_seed(100 + rank)
    return torch.randn(5, 6, generator=g), torch.randint(0, 3, (5,), generator=g)


def _worker(rank, world, port, defer, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from istgcn import dp
    torch.manual_seed(0)
    model = Tiny()
    if rank == 1:                                        # replicas start different: broadcast fixes it
        with torch.no_grad():
            for p in model.parameters():
                p.add_(1.0)
    dp.broadcast_state(model)
    buckets = dp.GradBuckets(list(model.named_parameters()), bucket_bytes=256)
    buckets.defer = defer
    assert any('branch.bn' in n for n in buckets.skipped)
    assert len(buckets.buckets) >= 2
    first = [n for n, _ in buckets.buckets[0]['params']]
    assert first[0].startswith('c.')                     # reverse registration order
    for step in range(2):
        x, y = _data(rank)
        buckets.zero()
        loss = nn.functional.cross_entropy(model(x + step), y)
        loss.backward()
        if not defer:
            assert all(b['launched'] for b in buckets.buckets)   # launched from the hooks
        buckets.finish()
    grads = {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}
    if rank == 0:
        torch.save({'grads': grads, 'state': model.state_dict()}, out)
    dist.destroy_process_group()


def _replay_worker(rank, world, port, out):
    """What a CUDA-graph replay looks like to the host: the flat gradient buffers are refilled by
    the device, neither the hooks nor zero() run, only finish() is called -- step after step."""
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from istgcn import dp
    torch.manual_seed(0)
    model = Tiny()
    before = {n: p.detach().clone() for n, p in model.named_parameters()}
    buckets = dp.GradBuckets(list(model.named_parameters()), bucket_bytes=256, flatten_params=True)
    for n, p in model.named_parameters():                # flattening keeps values, 16-byte alignment
        assert torch.equal(p.detach(), before[n])
        if not dp.is_unused(n):
            assert p.data_ptr() % 16 == 0 and p.grad.data_ptr() % 16 == 0
    x, y = _data(rank)
    buckets.zero()
    nn.functional.cross_entropy(model(x), y).backward()
    buckets.finish()
    results = []
    for step in range(3):                                # three "replays"
        for b in buckets.buckets:
            b['flat'].fill_(float(rank + 1 + step))      # rank-local gradient written by the device
        buckets.finish()
        results.append([b['flat'].clone() for b in buckets.buckets])
    if rank == 0:
        torch.save(results, out)
    dist.destroy_process_group()


def test_finish_without_zero_still_reduces(tmp_path):
    """after the first finish() every later finish() skipped the all-reduce."""
    out = str(tmp_path / 'replay.pt')
    mp.spawn(_replay_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    for step, flats in enumerate(torch.load(out)):
        for flat in flats:
            assert torch.all(flat == 1.5 + step), (step, flat[:4])


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.parametrize('defer', [False, True])
def test_gradient_buckets_average_over_two_ranks(tmp_path, defer):
    out = str(tmp_path / 'rank0.pt')
    mp.spawn(_worker, args=(2, _free_port(), defer, out), nprocs=2, join=True)
    got = torch.load(out)
    torch.manual_seed(0)
    ref = Tiny()
    ref.load_state_dict(got['state'])
    acc = {}
    for rank in range(2):
        x, y = _data(rank)
        ref.zero_grad()
        nn.functional.cross_entropy(ref(x + 1), y).backward()
        for n, p in ref.named_parameters():
            if p.grad is not None:
                acc[n] = acc.get(n, 0) + p.grad / 2
    assert set(acc) == set(got['grads'])
    for n in acc:
        torch.testing.assert_close(got['grads'][n], acc[n], rtol=1e-5, atol=1e-6)


def test_trainer_helpers():
    from istgcn import trainer
    m = nn.Sequential(nn.Conv2d(3, 4, 1), nn.BatchNorm2d(4), nn.Conv1d(3, 2, 1), nn.Linear(2, 2))
    before = m[3].weight.clone()
    torch.manual_seed(0)
    m.apply(trainer.weights_init)
    assert m[0].bias.abs().max() == 0 and m[1].bias.abs().max() == 0
    assert abs(m[1].weight.mean().item() - 1.0) < 0.1 and m[0].weight.std() < 0.05
    assert torch.equal(m[3].weight, before)             # nn.Linear is left untouched
    opt = torch.optim.SGD(m.parameters(), lr=0.1)
    assert trainer.adjust_lr(opt, 0.1, [20, 40], 25) == pytest.approx(0.01)
    assert opt.param_groups[0]['lr'] == pytest.approx(0.01)
    assert trainer.adjust_lr(opt, 0.1, [], 25) == pytest.approx(0.1)
