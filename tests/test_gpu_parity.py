"""GPU parity tests (``-m gpu``): every call goes through the C ABI of libistgcn_b200.so and is
compared with the CPU oracle (oracle/model_ref.py, fp64 where stated) or with the golden
fixtures generated from the reference.

Tolerances (BASELINE.json north_star): <= 1e-4 relative in the fp32-grade mode ('3xtf32'),
<= 2e-2 in the fast tensor-core mode ('tf32').  "Relative" = max |a - b| / max |b| per tensor.
"""
import importlib.util
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

TOL = {'3xtf32': 1e-4, 'tf32': 2e-2}
TOL_GRAD = {'3xtf32': 5e-4, 'tf32': 2e-2}


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


@pytest.fixture(scope='module')
def env():
    import istgcn
    from istgcn import _lib
    _lib.load()
    return istgcn


def _mg():
    spec = importlib.util.spec_from_file_location(
        'make_golden', os.path.join(os.path.dirname(__file__), 'golden', 'make_golden.py'))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    return mg


# ----------------------------------------------------------------------------- graph conv op
@pytest.mark.parametrize('math', ['3xtf32', 'tf32'])
@pytest.mark.parametrize('layout,strategy,cin,cout,nm,t', [
    ('ntu-rgb+d_sym', 'spatial_3_sym', 64, 64, 3, 7),      # partial last tile (21 frames, 5/tile)
    ('ntu-rgb+d_sym', 'spatial_3_sym', 3, 64, 2, 9),       # block 0: Cin = 3
    ('ntu-rgb+d', 'spatial', 64, 128, 2, 5),               # K = 3, single A
    ('openpose_sym', 'spatial_3_sym', 128, 128, 2, 8),     # V = 18, 7 frames / tile
    ('ntu-rgb+d_sym', 'spatial_sym', 128, 256, 1, 6),
    ('ntu-rgb+d_sym', 'spatial_3_sym', 256, 256, 1, 5),
])
def test_graph_conv_op(env, math, layout, strategy, cin, cout, nm, t):
    from net.utils.graph import Graph
    from net.utils.inceptionv2_gcn import Inception2
    from net.utils.tgcn import ConvTemporalGraphical
    from oracle import model_ref
    dev = torch.device('cuda')
    g = Graph(layout, strategy)
    gen = torch.Generator().manual_seed(cin * 1000 + cout)
    K, V = g.A.shape[0], g.A.shape[1]
    stacks = [torch.tensor(getattr(g, n), dtype=torch.float32) for n in ('A', 'A2', 'A3')
              if hasattr(g, n)]
    adjs = [(a * (1 + 0.3 * torch.randn(a.shape, generator=gen))).requires_grad_(True)
            for a in stacks]
    x = torch.randn(nm, cin, t, V, generator=gen, requires_grad=True)
    mod = (Inception2 if len(adjs) == 3 else ConvTemporalGraphical)(cin, cout, K)
    conv = mod.branch.conv if len(adjs) == 3 else mod.conv
    with torch.no_grad():
        conv.weight.normal_(0, 0.1, generator=gen)
        conv.bias.normal_(0, 0.1, generator=gen)
    w64 = conv.weight.detach().double().requires_grad_(True)
    b64 = conv.bias.detach().double().requires_grad_(True)
    x64 = x.detach().double().requires_grad_(True)
    a64 = [a.detach().double().requires_grad_(True) for a in adjs]
    ref = model_ref.graph_conv(x64, w64, b64, a64)
    gout = torch.randn(ref.shape, generator=gen)
    ref.backward(gout.double())

    mod = mod.to(dev)
    xg = x.detach().to(dev).requires_grad_(True)
    ag = [a.detach().to(dev).requires_grad_(True) for a in adjs]
    old = env.set_math(math)
    try:
        out = mod(xg, *ag)[0]
        out.backward(gout.to(dev))
    finally:
        env.set_math(old)
    tol = TOL[math]
    assert rel(out, ref) < tol
    assert rel(xg.grad, x64.grad) < tol
    assert rel(conv.weight.grad, w64.grad) < tol
    assert rel(conv.bias.grad, b64.grad) < tol
    for mine, theirs in zip(ag, a64):
        if theirs.grad.abs().max() > 0:
            assert rel(mine.grad, theirs.grad) < tol


# ----------------------------------------------------------------------------- data_bn
def test_data_bn_layout_and_stats(env):
    from istgcn import ops
    dev = torch.device('cuda')
    gen = torch.Generator().manual_seed(3)
    N, C, T, V, M = 3, 3, 11, 25, 2
    x = torch.randn(N, C, T, V, M, generator=gen) * 2 + 0.5
    bn = torch.nn.BatchNorm1d(V * C)
    with torch.no_grad():
        bn.weight.normal_(1, 0.2, generator=gen)
        bn.bias.normal_(0, 0.2, generator=gen)
    ref_bn = torch.nn.BatchNorm1d(V * C)
    ref_bn.load_state_dict(bn.state_dict())
    xr = x.permute(0, 4, 3, 1, 2).contiguous().view(N * M, V * C, T)
    yr = ref_bn(xr).view(N, M, V, C, T).permute(0, 1, 4, 2, 3).contiguous().view(N * M, T, V, C)
    gy = torch.randn(yr.shape, generator=gen)
    yr.backward(gy)
    bn = bn.to(dev)
    y = ops.DataBN.apply(x.to(dev), bn.weight, bn.bias, ops.BNState(bn), True)
    y.backward(gy.to(dev))
    assert rel(y, yr) < 1e-5
    assert rel(bn.running_mean, ref_bn.running_mean) < 1e-5
    assert rel(bn.running_var, ref_bn.running_var) < 1e-5
    assert rel(bn.weight.grad, ref_bn.weight.grad) < 1e-4
    assert rel(bn.bias.grad, ref_bn.bias.grad) < 1e-4
    bn.eval(); ref_bn.eval()
    ye = ops.DataBN.apply(x.to(dev), bn.weight, bn.bias, ops.BNState(bn), False)
    yre = ref_bn(xr).view(N, M, V, C, T).permute(0, 1, 4, 2, 3).contiguous().view(N * M, T, V, C)
    assert rel(ye, yre) < 1e-5


# ----------------------------------------------------------------------------- one block
def _block_state(blk, prefix):
    return {prefix + k: v.detach().cpu().clone() for k, v in blk.state_dict().items()}


@pytest.mark.parametrize('math', ['3xtf32', 'tf32'])
@pytest.mark.parametrize('cin,cout,stride,residual,t', [
    (64, 64, 1, True, 23),        # identity residual, T not a multiple of the tile
    (64, 128, 2, True, 20),       # strided conv + BN residual
    (128, 256, 2, True, 15),      # odd T with stride 2 (T_out = 8)
    (3, 64, 1, False, 12),        # block 0
    (256, 256, 1, True, 9),
])
def test_block_vs_oracle(env, math, cin, cout, stride, residual, t):
    from net.ist_gcn import st_gcn
    from net.utils.graph import Graph
    from oracle import model_ref
    dev = torch.device('cuda')
    gen = torch.Generator().manual_seed(cin + cout + stride)
    g = Graph('ntu-rgb+d_sym', 'spatial_3_sym')
    K, V = 4, 25
    blk = st_gcn(cin, cout, (9, K), stride, residual=residual)
    with torch.no_grad():
        for m in blk.modules():
            if isinstance(m, torch.nn.Conv2d):
                m.weight.normal_(0, 0.08, generator=gen)
                m.bias.normal_(0, 0.05, generator=gen)
            elif isinstance(m, torch.nn.BatchNorm2d):
                m.weight.normal_(1, 0.1, generator=gen)
                m.bias.normal_(0, 0.1, generator=gen)
    state = _block_state(blk, 'b.')
    nm = 3
    x = torch.randn(nm, cin, t, V, generator=gen)
    adjs = [torch.tensor(getattr(g, n), dtype=torch.float32) *
            (1 + 0.2 * torch.randn(K, V, V, generator=gen)) for n in ('A', 'A2', 'A3')]
    m_imp = 1 + 0.3 * torch.randn(3, generator=gen)

    # oracle in float64
    st64 = {k: (v.double().requires_grad_(True) if v.is_floating_point() and 'running' not in k
                else (v.double() if v.is_floating_point() else v)) for k, v in state.items()}
    x64 = x.double().requires_grad_(True)
    a64 = [a.double().requires_grad_(True) for a in adjs]
    m64 = m_imp.double().requires_grad_(True)
    upd = {}
    ref = model_ref.block_forward(st64, 'b.', 'ist_gcn', x64, a64, m64, (cin, cout, stride, residual),
                                  True, 0.0, upd)
    gout = torch.randn(ref.shape, generator=gen)
    ref.backward(gout.double())

    blk = blk.to(dev).train()
    xg = x.to(dev).requires_grad_(True)
    ag = [a.to(dev).requires_grad_(True) for a in adjs]
    mg_ = m_imp.to(dev).requires_grad_(True)
    old = env.set_math(math)
    try:
        out = blk(xg, ag[0], ag[1], ag[2], mg_)[0]
        out.backward(gout.to(dev))
    finally:
        env.set_math(old)
    tol, tolg = TOL[math], TOL_GRAD[math]
    assert rel(out, ref) < tol, 'block output'
    errs = {}
    if cin > 3 or True:
        errs['x'] = rel(xg.grad, x64.grad)
    for name, p in blk.named_parameters():
        r = st64['b.' + name].grad
        if r is None:
            assert p.grad is None or p.grad.abs().max() == 0 or 'branch.bn' in name, name
            continue
        errs[name] = rel(p.grad, r)
    for i in range(3):
        if a64[i].grad.abs().max() > 0:
            errs['A%d' % i] = rel(ag[i].grad, a64[i].grad)
    errs['m_imp'] = rel(mg_.grad, m64.grad)
    bad = {k: v for k, v in errs.items() if not v < tolg}
    assert not bad, bad
    # running statistics of every BatchNorm that ran
    after = blk.state_dict()
    for k, v in upd.items():
        key = k[len('b.'):]
        if key.endswith('num_batches_tracked'):
            assert int(after[key]) == int(v)
        else:
            assert rel(after[key], v) < 1e-4, key


# ----------------------------------------------------------------------------- whole network
def _load_case(name):
    from net.utils.graph import Graph
    mg = _mg()
    g_args, num_class, shape = mg.MODEL_CASES[name]
    graph = Graph(**g_args)
    state = mg.case_state(name, graph)
    x, label = mg.case_inputs(name, shape, num_class)
    return mg, g_args, num_class, shape, state, x, label


@pytest.mark.parametrize('math', ['3xtf32', 'tf32'])
@pytest.mark.parametrize('name', ['ist_gcn', 'ist_gcn_kinetics', 'st_gcn_mstcn_1x1'])
def test_model_vs_golden_and_oracle(env, math, name, golden_dir):
    """Logits, loss and EVERY parameter gradient of a training step vs the fixture generated
    from the reference's own modules and vs the live oracle."""
    import net.ist_gcn
    import net.st_gcn_mstcn_1x1
    from oracle import model_ref
    mg, g_args, num_class, shape, state, x, label = _load_case(name)
    arch = name.replace('_kinetics', '')
    fix = np.load(os.path.join(golden_dir, 'model_%s.npz' % name))
    assert mg.state_digest(state) == str(fix['state_sha256'])
    cls = net.ist_gcn.Model if arch == 'ist_gcn' else net.st_gcn_mstcn_1x1.Model
    model = cls(shape[1], num_class, g_args, True)
    assert list(model.state_dict().keys()) == list(state.keys())
    model.load_state_dict(state, strict=True)
    dev = torch.device('cuda')
    model = model.to(dev)
    old = env.set_math(math)
    try:
        model.eval()
        with torch.no_grad():
            ev = model(x.to(dev))
        model.train()
        logits = model(x.to(dev))
        loss = F.cross_entropy(logits, label.to(dev))
        loss.backward()
    finally:
        env.set_math(old)
    tol, tolg = TOL[math], TOL_GRAD[math]
    assert rel(ev, torch.from_numpy(fix['logits_eval'])) < tol
    assert rel(logits, torch.from_numpy(fix['logits_train'])) < tol
    assert abs(loss.item() - float(fix['loss'])) < tol * abs(float(fix['loss']))
    assert rel(model.data_bn.running_mean, torch.from_numpy(fix['data_bn.running_mean'])) < 1e-5
    assert rel(model.data_bn.running_var, torch.from_numpy(fix['data_bn.running_var'])) < 1e-5
    params = dict(model.named_parameters())
    names = [str(s) for s in fix['grad_names']]
    assert sorted(k for k, p in params.items() if p.grad is not None) == sorted(names)
    bad = {}
    for k in names:
        ref = fix['grad|' + k]
        mine = mg.probe(params[k].grad.cpu())
        # probes: [L2 norm, sum, 48 samples]; compare samples and norm relative to the norm
        err = np.abs(mine[2:] - ref[2:]).max() / max(np.abs(ref[2:]).max(), 1e-30)
        nerr = abs(mine[0] - ref[0]) / max(ref[0], 1e-30)
        if not (err < 5 * tolg and nerr < tolg):
            bad[k] = (err, nerr)
    assert not bad, bad


def test_dropout_mask_is_what_the_kernels_use(env):
    """Training step with dropout=0.5: read the keep-masks back through the C ABI, inject them
    into the oracle and require the same logits / gradients."""
    import net.ist_gcn
    from istgcn import ops
    from oracle import model_ref
    mg, g_args, num_class, shape, state, x, label = _load_case('ist_gcn')
    p = 0.5
    model = net.ist_gcn.Model(shape[1], num_class, g_args, True, dropout=p)
    model.load_state_dict(state, strict=True)
    dev = torch.device('cuda')
    model = model.to(dev).train()
    old = env.set_math('3xtf32')
    try:
        logits = model(x.to(dev))
        F.cross_entropy(logits, label.to(dev)).backward()
    finally:
        env.set_math(old)
    masks = {}
    N, C, T, V, M = shape
    t = T
    for i, blk in enumerate(model.st_gcn_networks):
        cin, cout, stride = blk._io
        t = (t - 1) // stride + 1
        if i == 0:
            continue
        numel = N * M * t * V * cout
        m = ops.dropout_mask(numel, p, blk.last_seed, dev).view(N * M, t, V, cout)
        frac = m.float().mean().item()
        assert abs(frac - (1 - p)) < 0.02, frac
        masks['st_gcn_networks.%d.' % i] = m.permute(0, 3, 1, 2).cpu()
    leaves = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and 'running' not in k
                  and k not in ('A', 'A2', 'A3') else v) for k, v in state.items()}
    ref = model_ref.forward(leaves, x, 'ist_gcn', training=True, dropout=p, masks=masks)
    F.cross_entropy(ref, label).backward()
    assert rel(logits, ref) < 1e-4
    worst = max(rel(prm.grad, leaves[k].grad) for k, prm in model.named_parameters()
                if leaves[k].grad is not None)
    assert worst < 2e-3, worst


def test_cpu_input_raises(env):
    import net.ist_gcn
    model = net.ist_gcn.Model(3, 60, dict(layout='ntu-rgb+d_sym', strategy='spatial_3_sym'), True)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        model(torch.zeros(1, 3, 8, 25, 2))
