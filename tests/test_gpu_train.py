"""GPU tests of the BENCHMARKED path (``-m gpu``): the training step that bench.py times
(istgcn.trainer.Trainer.step, eager and CUDA-graph replay, flat buckets + one-kernel SGD), the
BASELINE.json layer shapes (NM = 128, T = 300 / 150 / 75) in the fast 'tf32' mode, the CUDA
``extract_feature`` and -- with two GPUs -- the NCCL data-parallel step.

reference lines: processor/recognition.py:249-298 (iteration), :152-159 (optimiser),
net/st_gcnold.py:98-120 (extract_feature), processor/my_io.py:77-87 (DataParallel wrap).

The checker is the oracle (oracle/model_ref.py) evaluated in fp64 -- on the GPU for the full-size
cases, where the CPU would need minutes.  Bars: '3xtf32' 1e-4 on forward quantities, 'tf32' 2e-3 on
logits / loss and 5e-3 per operator (5x the observed error; the mode's budget in BASELINE.json is
2e-2).  Gradients in 'tf32' are calibrated against what stock PyTorch gets with TF32 enabled
(``torch.backends.cudnn.allow_tf32`` + ``cuda.matmul.allow_tf32``) on the same graph: relative
L2 error vs fp64 per tensor <= max(floor, 4 x PyTorch-TF32's own error for that tensor).
"""
import importlib.util
import os
import socket

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

TOL_FWD = {'3xtf32': 1e-4, 'tf32': 2e-3}


def rel(a, b, floor=1e-30):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(floor)).item()


def rel_l2(a, b, floor=1e-30):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(floor)).item()


def _mg():
    spec = importlib.util.spec_from_file_location(
        'make_golden', os.path.join(os.path.dirname(__file__), 'golden', 'make_golden.py'))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    return mg


def _case(name):
    from net.utils.graph import Graph
    mg = _mg()
    g_args, num_class, shape = mg.MODEL_CASES[name]
    state = mg.case_state(name, Graph(**g_args))
    x, label = mg.case_inputs(name, shape, num_class)
    return mg, g_args, num_class, shape, state, x, label


@pytest.fixture(scope='module')
def env():
    import istgcn
    from istgcn import _lib
    _lib.load()
    return istgcn


def _grad_names(state):
    return [k for k, v in state.items() if v.is_floating_point() and 'running' not in k
            and k not in ('A', 'A2', 'A3') and '.gcn.branch.bn.' not in k and '.linear.' not in k]


class OracleTrainer(object):
    """recognition.py:249-298 on the oracle: forward, cross-entropy, backward, SGD(nesterov)."""

    def __init__(self, state, arch, lr, dtype=torch.float64, device='cuda'):
        self.state = {k: (v.to(device=device, dtype=dtype) if v.is_floating_point() else v.to(device))
                      for k, v in state.items()}
        self.arch, self.lr = arch, lr
        self.names = _grad_names(state)
        self.bufs = [None] * len(self.names)

    def grads(self, x, label, dropout=0.0, masks=None):
        from oracle import model_ref
        for k in self.names:
            self.state[k] = self.state[k].detach().requires_grad_(True)
        upd = {}
        out = model_ref.forward(self.state, x.to(self.state['A']), self.arch, training=True,
                                dropout=dropout, update=upd, masks=masks)
        loss = F.cross_entropy(out, label.to(out.device))
        grads = torch.autograd.grad(loss, [self.state[k] for k in self.names])
        for k in self.names:
            self.state[k] = self.state[k].detach()
        return loss.detach(), grads, upd

    def step(self, x, label):
        from oracle import model_ref
        loss, grads, upd = self.grads(x, label)
        with torch.no_grad():
            model_ref.sgd_nesterov_step([self.state[k] for k in self.names], list(grads), self.bufs,
                                        self.lr)
            self.state.update(upd)
        return loss


def _model(arch, shape, num_class, g_args, state, dropout=0.0):
    import importlib
    cls = importlib.import_module('net.' + arch).Model
    kw = {'dropout': dropout} if dropout else {}
    model = cls(shape[1], num_class, g_args, True, **kw)
    model.load_state_dict(state, strict=True)
    return model.cuda()


# ----------------------------------------------------------------------------- training step
def _momentum_views(tr):
    """name -> the momentum buffer of that parameter (a view into FlatSGD's flat state)."""
    out = {}
    for b, buf in zip(tr.buckets.buckets, tr.optimizer.state):
        for n, prm in b['params']:
            off = prm.grad.storage_offset()
            out[n] = buf[off:off + prm.numel()].view_as(prm)
    return out


def _expected_step(before, bufs, arch, lr, x, label, dtype, tf32_flags=(False, False)):
    """One oracle iteration (recognition.py:273-289) from the given pre-step state:
    -> loss, {name: parameter after the step}, {BatchNorm buffers after the step}."""
    from oracle import model_ref
    old = _tf32_torch(tf32_flags)
    try:
        ora = OracleTrainer(before, arch, lr, dtype=dtype)
        loss, grads, upd = ora.grads(x, label)
        params = [ora.state[k].clone() for k in ora.names]
        model_ref.sgd_nesterov_step(params, list(grads), [bufs[k].to(params[0]).clone() for k in ora.names], lr)
    finally:
        _tf32_torch(old)
    return loss.item(), dict(zip(ora.names, params)), upd


@pytest.mark.parametrize('math', ['3xtf32', 'tf32'])
@pytest.mark.parametrize('mode', ['eager', 'graph'])
def test_trainer_step_vs_oracle(env, math, mode):
    """Five iterations of Trainer.step -- eager launches, or CUDA-graph capture + replay (steps
    3..5 are replays) -- each checked against ONE oracle iteration in fp64 started from the
    trainer's own pre-step state (parameters, momentum, running statistics): loss, parameter
    update, BatchNorm running statistics.  (Comparing whole trajectories instead measures chaos:
    by step 4 stock fp32 PyTorch is 1e-2 away from fp64 on the loss.)

    The update is bounded relative to stock PyTorch doing the same single step on this GPU in fp32
    (for '3xtf32') or with TF32 enabled (for 'tf32'): relative L2 distance to the fp64 update
    <= max(floor, 4 x PyTorch's own distance)."""
    from istgcn import trainer
    mg, g_args, num_class, shape, state, x, label = _case('ist_gcn')
    lr, steps = 0.05, 5
    gen = torch.Generator().manual_seed(9)
    xs = [x] + [torch.randn(shape, generator=gen) for _ in range(steps - 1)]
    ys = [label] + [torch.randint(0, num_class, (shape[0],), generator=gen) for _ in range(steps - 1)]
    tol = TOL_FWD[math]
    old = env.set_math(math)
    try:
        model = _model('ist_gcn', shape, num_class, g_args, state)
        tr = trainer.Trainer(model, base_lr=lr, use_graph=(mode == 'graph'))
        mom = _momentum_views(tr)
        for i in range(steps):
            before = {k: v.detach().clone().cpu() for k, v in model.state_dict().items()}
            bufs = {k: v.detach().clone().double() for k, v in mom.items()}
            loss = tr.step(xs[i].cuda(), ys[i].cuda()).item()
            after = {k: v.detach().clone().cpu() for k, v in model.state_dict().items()}
            ref_loss, ref_p, ref_upd = _expected_step(before, bufs, 'ist_gcn', lr, xs[i], ys[i], torch.float64)
            _, cal_p, _ = _expected_step(before, bufs, 'ist_gcn', lr, xs[i], ys[i], torch.float32,
                                         (True, True) if math == 'tf32' else (False, False))
            assert abs(loss - ref_loss) < tol * abs(ref_loss), (i, loss, ref_loss)
            names = list(ref_p)
            dm = torch.cat([(after[k].double() - before[k].double()).reshape(-1) for k in names])
            dr = torch.cat([(ref_p[k].cpu() - before[k].double()).reshape(-1) for k in names])
            dc = torch.cat([(cal_p[k].double().cpu() - before[k].double()).reshape(-1) for k in names])
            e_mine, e_cal = rel_l2(dm, dr), rel_l2(dc, dr)
            print('trainer %s %s step %d: loss %.6f (oracle %.6f) update rel-L2 %.2e (pytorch %.2e)' % (
                math, mode, i, loss, ref_loss, e_mine, e_cal))
            assert e_mine < max(2e-3 if math == '3xtf32' else 2e-2, 4 * e_cal), (i, e_mine, e_cal)
            for k, v in ref_upd.items():
                if '.gcn.branch.bn.' in k:
                    continue
                if k.endswith('num_batches_tracked'):
                    assert int(after[k]) == int(v), k
                else:
                    assert rel(after[k], v) < 2.5 * tol, (i, k)
        if mode == 'graph':
            assert len(tr._graphs) == 1, 'steps 3.. must have been graph replays'
    finally:
        env.set_math(old)


def test_graph_replay_with_dropout_draws_fresh_masks_and_matches_oracle(env):
    """CUDA-graph replays with dropout 0.5: every replay draws a new mask (device-resident step
    counter); the masks of the last replay, read back through the C ABI and injected into the
    oracle together with the parameters / momentum from before that step, reproduce its loss and
    its parameter update."""
    from istgcn import ops, trainer
    mg, g_args, num_class, shape, state, x, label = _case('ist_gcn')
    p, lr = 0.5, 0.05
    model = _model('ist_gcn', shape, num_class, g_args, state, dropout=p)
    old = env.set_math('3xtf32')
    try:
        tr = trainer.Trainer(model, base_lr=lr, use_graph=True)
        xd, yd = x.cuda(), label.cuda()
        for _ in range(3):
            tr.step(xd, yd)
        assert len(tr._graphs) == 1
        losses = [tr.step(xd, yd).item() for _ in range(2)]
        assert losses[0] != losses[1]
        before = {k: v.detach().clone().cpu() for k, v in model.state_dict().items()}
        bufs = {n: v.detach().clone().double() for n, v in _momentum_views(tr).items()}
        loss = tr.step(xd, yd).item()
    finally:
        env.set_math(old)
    N, C, T, V, M = shape
    masks, t = {}, T
    for i, blk in enumerate(model.st_gcn_networks):
        cin, cout, stride = blk._io
        t = (t - 1) // stride + 1
        if i == 0:
            continue
        m = ops.dropout_mask(N * M * t * V * cout, p, blk.last_seed, xd.device).view(N * M, t, V, cout)
        masks['st_gcn_networks.%d.' % i] = m.permute(0, 3, 1, 2)
    ora = OracleTrainer(before, 'ist_gcn', lr)
    ref_loss, grads, _ = ora.grads(x, label, dropout=p, masks=masks)
    assert abs(loss - ref_loss.item()) < 1e-4 * abs(ref_loss.item())
    from oracle import model_ref
    params = [ora.state[k].clone() for k in ora.names]
    model_ref.sgd_nesterov_step(params, list(grads), [bufs[k].cuda() for k in ora.names], lr)
    after = model.state_dict()
    dm = torch.cat([(after[k].double().cpu() - before[k].double()).reshape(-1) for k in ora.names])
    dr = torch.cat([(pr.cpu() - before[k].double()).reshape(-1) for k, pr in zip(ora.names, params)])
    assert rel_l2(dm, dr) < 2e-2, rel_l2(dm, dr)


def test_lr_schedule_reaches_captured_graph(env):
    """The optimiser kernel reads the learning rate from device memory: set_lr(0) after capture
    must freeze the parameters on the next replay (a baked-in rate would keep moving them)."""
    from istgcn import trainer
    mg, g_args, num_class, shape, state, x, label = _case('ist_gcn')
    model = _model('ist_gcn', shape, num_class, g_args, state)
    tr = trainer.Trainer(model, base_lr=0.05, use_graph=True, momentum=0.0, weight_decay=0.0)
    xd, yd = x.cuda(), label.cuda()
    for _ in range(3):
        tr.step(xd, yd)
    assert len(tr._graphs) == 1
    snap = [p.detach().clone() for p in model.parameters()]
    tr.set_lr(0.0)
    tr.step(xd, yd)
    assert all(torch.equal(a, b.detach()) for a, b in zip(snap, model.parameters()))
    tr.set_lr(0.05)
    tr.step(xd, yd)
    assert any(not torch.equal(a, b.detach()) for a, b in zip(snap, model.parameters()))


def test_sgd_kernel_vs_torch(env):
    from istgcn._lib import call, i64
    gen = torch.Generator().manual_seed(1)
    n = 4099
    p = torch.randn(n + 1, generator=gen).cuda()[:n]
    g = torch.randn(n + 1, generator=gen).cuda()[:n]
    ref_p = torch.nn.Parameter(p.clone().double())
    opt = torch.optim.SGD([ref_p], lr=0.1, momentum=0.9, nesterov=True, weight_decay=1e-4)
    buf = torch.zeros_like(p)
    lr = torch.tensor([0.1], device='cuda')
    for step in range(3):
        ref_p.grad = (g * (step + 1) * 0.5).double()          # averaged gradient of two "ranks"
        opt.step()
        call('sgd_step', p, (g * (step + 1)).contiguous(), buf, i64(n), lr, 0.9, 1e-4, 1, 0.5)
        assert rel(p, ref_p) < 1e-6


# ----------------------------------------------------------------------------- full-size layers
def _tf32_torch(flag):
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = flag
    return old


@pytest.mark.parametrize('cin,cout,stride,t,layout,nm', [
    (64, 64, 1, 300, 'ntu-rgb+d_sym', 128), (64, 128, 2, 300, 'ntu-rgb+d_sym', 128),
    (128, 128, 1, 150, 'ntu-rgb+d_sym', 128), (128, 256, 2, 150, 'ntu-rgb+d_sym', 128),
    (256, 256, 1, 75, 'ntu-rgb+d_sym', 128),
    (128, 128, 1, 150, 'openpose_sym', 512)])       # BASELINE cfg 3: Kinetics-skeleton, batch 256, V = 18
def test_block_at_baseline_shape_tf32(env, cin, cout, stride, t, layout, nm):
    """One IST-GCN block at the BASELINE.json cfg-2 layer shape (batch 64 -> NM = 128 person
    sequences, V = 25: 38 400 / 19 200 / 9 600 frames, every persistent CTA walks many tiles) in
    the benchmarked 'tf32' mode vs the oracle in fp64 ON THE GPU; gradients calibrated against
    stock PyTorch with TF32 enabled."""
    from net.ist_gcn import st_gcn
    from net.utils.graph import Graph
    from oracle import model_ref
    dev = torch.device('cuda')
    gen = torch.Generator().manual_seed(cin + cout + stride)
    g = Graph(layout, 'spatial_3_sym')
    K, V = 4, g.A.shape[1]
    blk = st_gcn(cin, cout, (9, K), stride, residual=True)
    with torch.no_grad():
        for m in blk.modules():
            if isinstance(m, torch.nn.Conv2d):
                m.weight.normal_(0, 0.08, generator=gen)
                m.bias.normal_(0, 0.05, generator=gen)
            elif isinstance(m, torch.nn.BatchNorm2d):
                m.weight.normal_(1, 0.1, generator=gen)
                m.bias.normal_(0, 0.1, generator=gen)
    state = {'b.' + k: v.detach().clone() for k, v in blk.state_dict().items()}
    x = torch.randn(nm, cin, t, V, generator=gen)
    adjs = [torch.tensor(getattr(g, n), dtype=torch.float32) *
            (1 + 0.2 * torch.randn(K, V, V, generator=gen)) for n in ('A', 'A2', 'A3')]
    m_imp = 1 + 0.3 * torch.randn(3, generator=gen)
    tout = (t - 1) // stride + 1
    gout = torch.randn(nm, cout, tout, V, generator=gen)

    def oracle(dtype):
        st = {k: (v.to(dev, dtype).requires_grad_(True) if v.is_floating_point() and 'running' not in k
                  else (v.to(dev, dtype) if v.is_floating_point() else v.to(dev))) for k, v in state.items()}
        xx = x.to(dev, dtype).requires_grad_(True)
        aa = [a.to(dev, dtype).requires_grad_(True) for a in adjs]
        mm = m_imp.to(dev, dtype).requires_grad_(True)
        out = model_ref.block_forward(st, 'b.', 'ist_gcn', xx, aa, mm, (cin, cout, stride, True), True, 0.0, {})
        out.backward(gout.to(dev, dtype))
        grads = {'x': xx.grad, 'm_imp': mm.grad}
        grads.update({k: v.grad for k, v in st.items() if getattr(v, 'grad', None) is not None})
        for i in range(3):
            grads['A%d' % i] = aa[i].grad
        res = out.detach().double().cpu(), {k: v.detach().double().cpu() for k, v in grads.items()}
        del st, xx, aa, mm, out, grads
        torch.cuda.empty_cache()
        return res

    ref, g64 = oracle(torch.float64)
    old_flags = _tf32_torch((True, True))
    try:
        ref32, g32 = oracle(torch.float32)
    finally:
        _tf32_torch(old_flags)
    blk = blk.to(dev).train()
    xg = x.to(dev).requires_grad_(True)
    ag = [a.to(dev).requires_grad_(True) for a in adjs]
    mg_ = m_imp.to(dev).requires_grad_(True)
    old = env.set_math('tf32')
    try:
        out = blk(xg, ag[0], ag[1], ag[2], mg_)[0]
        out.backward(gout.to(dev))
    finally:
        env.set_math(old)
    torch.cuda.synchronize()
    e_out, e_out32 = rel(out, ref), rel(ref32, ref)
    print('block %d->%d s%d T%d: out rel %.2e (torch-tf32 %.2e)' % (cin, cout, stride, t, e_out, e_out32))
    assert e_out < 5e-3
    mine = {'x': xg.grad, 'm_imp': mg_.grad}
    mine.update({'b.' + k: p.grad for k, p in blk.named_parameters() if p.grad is not None})
    union = ((adjs[0] != 0) | (adjs[1] != 0) | (adjs[2] != 0)).double()
    for i in range(3):
        mine['A%d' % i] = ag[i].grad.double().cpu() * union
        g64['A%d' % i] = g64['A%d' % i] * union
        g32['A%d' % i] = g32['A%d' % i] * union
    gmax = max(v.abs().max().item() for k, v in g64.items() if k.startswith('b.'))
    bad = {}
    for k, v in sorted(mine.items()):
        if k not in g64:
            continue
        if k.startswith('b.') and g64[k].abs().max().item() < 1e-6 * gmax:
            # mathematically zero (a bias in front of a train-mode BatchNorm)
            assert v.detach().abs().max().item() < 2e-3 * gmax * (x.numel() / cin) ** 0.5, k
            continue
        e_mine, e_ref = rel_l2(v, g64[k]), rel_l2(g32[k], g64[k])
        print('   grad %-34s rel-L2 %.2e (torch-tf32 %.2e)' % (k, e_mine, e_ref))
        if not e_mine < max(5e-3, 4 * e_ref):
            bad[k] = (e_mine, e_ref)
    assert not bad, bad


# ----------------------------------------------------------------------------- extract_feature
@pytest.mark.parametrize('math', ['3xtf32', 'tf32'])
def test_extract_feature_vs_golden_and_oracle(env, math, golden_dir):
    """Model.extract_feature (net/st_gcnold.py:98-120) on the CUDA path: per-(t, v, m) logits and
    256-channel features vs the probes the reference's own st_gcnold produced and vs the oracle."""
    from oracle import model_ref
    mg, g_args, num_class, shape, state, x, label = _case('st_gcn')
    fix = np.load(os.path.join(golden_dir, 'model_st_gcn.npz'))
    model = _model('st_gcn', shape, num_class, g_args, state).eval()
    old = env.set_math(math)
    try:
        with torch.no_grad():
            out, feat = model.extract_feature(x.cuda())
    finally:
        env.set_math(old)
    N, C, T, V, M = shape
    assert tuple(out.shape) == (N, num_class, T // 4, V, M) and tuple(feat.shape) == (N, 256, T // 4, V, M)
    tol = TOL_FWD[math] * 2.5
    ro, rf = model_ref.extract_feature({k: v.double() if v.is_floating_point() else v for k, v in state.items()},
                                       x.double(), 'st_gcn')
    assert rel(out, ro) < tol and rel(feat, rf) < tol
    for mine, key in ((out, 'feat_out'), (feat, 'feat_feature')):
        probe, ref = mg.probe(mine.cpu()), fix[key]
        assert np.abs(probe[2:] - ref[2:]).max() < tol * np.abs(ref[2:]).max(), key
        assert abs(probe[0] - ref[0]) < tol * abs(ref[0]), key                  # L2 norm of the tensor
    # the Inception variants share the trunk: shapes + oracle (the reference's own method is broken
    # there, SURVEY.md App. E: it calls gcn(x, A) with the 2-argument signature)
    mg, g_args, num_class, shape, state, x, label = _case('ist_gcn')
    model = _model('ist_gcn', shape, num_class, g_args, state).eval()
    old = env.set_math(math)
    try:
        with torch.no_grad():
            out, feat = model.extract_feature(x.cuda())
    finally:
        env.set_math(old)
    ro, rf = model_ref.extract_feature({k: v.double() if v.is_floating_point() else v for k, v in state.items()},
                                       x.double(), 'ist_gcn')
    assert rel(out, ro) < tol and rel(feat, rf) < tol


# ----------------------------------------------------------------------------- data parallel (NCCL)
def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _dp_worker(rank, world, port, use_graph, out_path):
    os.environ['ISTGCN_GRAPH_COLLECTIVES'] = '1' if use_graph == 'collectives' else '0'
    use_graph = bool(use_graph)
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (os.path.join(root, 'ist-gcn_b200'), root):
        if p not in sys.path:
            sys.path.insert(0, p)
    import datetime
    import faulthandler
    import torch.distributed as dist
    faulthandler.dump_traceback_later(100, exit=True)            # a hung collective must not eat the GPU budget
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank),
                            timeout=datetime.timedelta(seconds=90))
    import istgcn
    from istgcn import dp, trainer
    istgcn.set_math('3xtf32')
    mg, g_args, num_class, shape, state, x, label = _case('ist_gcn')
    model = _model('ist_gcn', shape, num_class, g_args, state)
    if rank == 1:
        with torch.no_grad():
            for prm in model.parameters():
                prm.add_(0.5)                                   # broadcast_state must undo this
    dp.broadcast_state(model)
    tr = trainer.Trainer(model, base_lr=0.05, use_graph=use_graph, bucket_bytes=1 << 20)
    assert len(tr.buckets.buckets) >= 2
    gen = torch.Generator().manual_seed(100 + rank)
    xs = [torch.randn(shape, generator=gen) for _ in range(5)]
    ys = [torch.randint(0, num_class, (shape[0],), generator=gen) for _ in range(5)]
    # step 1 with the optimiser disabled: the averaged gradient itself
    tr._iteration(xs[0].cuda(), ys[0].cuda(), with_optimizer=False)
    tr.buckets.finish(average=True)
    grads = {n: prm.grad.detach().clone().cpu() for n, prm in model.named_parameters() if prm.grad is not None
             and not dp.is_unused(n)}
    model.load_state_dict(state, strict=True)                   # running statistics back to the start
    losses = [tr.step(xs[i].cuda(), ys[i].cuda()).item() for i in range(5)]
    if use_graph:
        assert len(tr._graphs) == 1
    equal = dp.replicas_equal(model)
    tr.close()                                                  # graphs with NCCL kernels go before the group
    if rank == 0:
        torch.save({'grads': grads, 'losses': losses, 'equal': equal,
                    'params': {n: prm.detach().cpu() for n, prm in model.named_parameters()}}, out_path)
    dist.destroy_process_group()
    faulthandler.cancel_dump_traceback_later()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')
def test_data_parallel_two_ranks_nccl(tmp_path):
    """Two NCCL ranks on the real net.ist_gcn: (1) the averaged gradient equals the mean of the
    per-shard oracle gradients (per-rank BatchNorm); (2) replicas stay
    bit-equal over five steps in eager mode, in CUDA-graph mode (graph = forward + backward, the
    all-reduces follow the replay) and with the all-reduces captured inside the graph
    (ISTGCN_GRAPH_COLLECTIVES=1); (3) all modes end at the same parameters."""
    import torch.multiprocessing as mp
    res = {}
    for use_graph in (False, True, 'collectives'):
        out = str(tmp_path / ('dp_%s.pt' % use_graph))
        mp.spawn(_dp_worker, args=(2, _free_port(), use_graph, out), nprocs=2, join=True)
        res[use_graph] = torch.load(out)
        assert res[use_graph]['equal'], 'replicas diverged (graph=%s)' % use_graph
    mg, g_args, num_class, shape, state, x, label = _case('ist_gcn')
    acc = None
    for rank in range(2):
        gen = torch.Generator().manual_seed(100 + rank)
        xr = torch.randn(shape, generator=gen)
        for _ in range(4):
            torch.randn(shape, generator=gen)
        yr = torch.randint(0, num_class, (shape[0],), generator=gen)
        ora = OracleTrainer(state, 'ist_gcn', 0.05)
        _, grads, _ = ora.grads(xr, yr)
        acc = [g / 2 for g in grads] if acc is None else [a + g / 2 for a, g in zip(acc, grads)]
    mine = torch.cat([res[False]['grads'][k].double().reshape(-1) for k in ora.names])
    ref = torch.cat([g.cpu().reshape(-1) for g in acc])
    assert rel_l2(mine, ref) < 2e-2, rel_l2(mine, ref)
    pe = torch.cat([v.double().reshape(-1) for v in res[False]['params'].values()])
    for mode in (True, 'collectives'):
        pg = torch.cat([v.double().reshape(-1) for v in res[mode]['params'].values()])
        assert rel_l2(pg, pe) < 1e-3, mode
        # (five steps at lr 0.05 amplify the atomics-order noise between the runs)
        assert max(abs(a - b) / abs(a) for a, b in zip(res[False]['losses'], res[mode]['losses'])) < 5e-3, mode


# ----------------------------------------------------------------------------- A/B switches
@pytest.mark.parametrize('switch', ['ISTGCN_PAIR_ASYNC=0', 'ISTGCN_DW_ASYNC=0', 'ISTGCN_BN_FOLD=0',
                                    'ISTGCN_PAIR_NB_MAX=128', 'ISTGCN_PAIR_BLOCKS=0', 'ISTGCN_PAIR_BLOCKS=1',
                                    'ISTGCN_SMALL_BWD_TC=0'])
def test_ab_switches_give_the_same_step(env, monkeypatch, switch):
    """The environment switches documented in DESIGN.md section 6 select alternative schedules of the
    SAME arithmetic (gradient stream on / off, BatchNorm bookkeeping folded into its consumers or not,
    item shapes of the pair-moment gradient): loss, every parameter gradient and the BatchNorm running
    statistics of one fast-mode training iteration must agree with the default path to rounding
    (bound: 10x the run-to-run noise of the default path, see below)."""
    mg, g_args, num_class, shape, state, x, label = _case('ist_gcn')
    env.set_math('tf32')
    x, label = x.cuda(), label.cuda()

    def run():
        torch.manual_seed(0)
        model = _model('ist_gcn', shape, num_class, g_args, state)
        model.train()
        # item plans are cached per pattern object: a fresh model re-plans under the current switches
        loss = F.cross_entropy(model(x), label)
        loss.backward()
        torch.cuda.synchronize()
        grads = {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}
        stats = {n: b.detach().clone() for n, b in model.named_buffers() if 'running' in n}
        return loss.item(), grads, stats

    loss0, g0, s0 = run()
    loss0b, g0b, s0b = run()
    key, val = switch.split('=')
    monkeypatch.setenv(key, val)
    loss1, g1, s1 = run()
    monkeypatch.delenv(key)
    # two runs of the DEFAULT path already differ in the last bits (atomic accumulation order of the
    # BatchNorm sums and weight gradients, amplified through ten blocks): the bound is 10x that noise,
    # with a floor (1e-2 on gradients) far below what a wrong schedule (a missed join, a stale
    # coefficient: errors of order one) would produce
    noise_l = abs(loss0b - loss0)
    assert abs(loss1 - loss0) <= max(10 * noise_l, 1e-3 * max(1.0, abs(loss0)))
    assert g0.keys() == g1.keys()
    # all gradients as one vector (the few-element importance parameters are sums with heavy cancellation
    # on this two-clip batch: their own run-to-run noise is percents), then the big tensors one by one
    flat = lambda g: torch.cat([g[n].flatten() for n in sorted(g)])              # noqa: E731
    noise = rel_l2(flat(g0b), flat(g0))
    assert rel_l2(flat(g1), flat(g0)) < max(10 * noise, 1e-2)
    for n in g0:
        if g0[n].numel() < 256:
            continue
        noise = rel_l2(g0b[n], g0[n], floor=1e-12)
        assert rel_l2(g1[n], g0[n], floor=1e-12) < max(20 * noise, 2e-2), n
    for n in s0:
        noise = rel_l2(s0b[n], s0[n], floor=1e-12)
        assert rel_l2(s1[n], s0[n], floor=1e-12) < max(10 * noise, 2e-3), n
