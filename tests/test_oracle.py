"""Pin the oracle (oracle/model_ref.py) to the reference: golden fixtures generated from the
reference's own modules (tests/golden/make_golden.py) and, when mounted, the live reference."""
import importlib.util
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from net.utils.graph import Graph
from oracle import model_ref, refload

_spec = importlib.util.spec_from_file_location(
    'make_golden', os.path.join(os.path.dirname(__file__), 'golden', 'make_golden.py'))
mg = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(mg)

CASES = sorted(mg.MODEL_CASES)


def _run_oracle(name, dtype=torch.float32):
    g_args, num_class, shape = mg.MODEL_CASES[name]
    arch = name.replace('_kinetics', '')
    graph = Graph(**g_args)
    state = mg.case_state(name, graph)
    x, label = mg.case_inputs(name, shape, num_class)
    return arch, state, x, label


@pytest.mark.parametrize('name', CASES)
def test_oracle_matches_golden(name, golden_dir):
    fix = np.load(os.path.join(golden_dir, 'model_%s.npz' % name))
    arch, state, x, label = _run_oracle(name)
    assert mg.state_digest(state) == str(fix['state_sha256']), 'seeded state differs from fixture'
    np.testing.assert_array_equal(x.numpy(), fix['x'])
    with torch.no_grad():
        ev = model_ref.forward(state, x, arch, training=False)
    np.testing.assert_allclose(ev.numpy(), fix['logits_eval'], rtol=1e-5, atol=1e-6)
    leaves = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and
                  k not in ('A', 'A2', 'A3') and 'running' not in k else v)
              for k, v in state.items()}
    upd = {}
    logits = model_ref.forward(leaves, x, arch, training=True, update=upd)
    loss = F.cross_entropy(logits, torch.as_tensor(label))
    loss.backward()
    np.testing.assert_allclose(logits.detach().numpy(), fix['logits_train'], rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(loss.item(), float(fix['loss']), rtol=1e-6)
    np.testing.assert_allclose(upd['data_bn.running_mean'].numpy(), fix['data_bn.running_mean'],
                               rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(upd['data_bn.running_var'].numpy(), fix['data_bn.running_var'],
                               rtol=1e-6, atol=1e-7)
    got = sorted(k for k, v in leaves.items() if v.requires_grad and v.grad is not None)
    assert got == sorted(str(s) for s in fix['grad_names'])
    for k in got:
        ref = fix['grad|' + k]
        mine = mg.probe(leaves[k].grad)
        scale = max(ref[0], 1e-12)                      # L2 norm of the reference gradient
        assert np.abs(mine - ref).max() <= 2e-5 * scale + 1e-9, k


def test_extract_feature_matches_golden(golden_dir):
    fix = np.load(os.path.join(golden_dir, 'model_st_gcn.npz'))
    arch, state, x, _ = _run_oracle('st_gcn')
    with torch.no_grad():
        out, feat = model_ref.extract_feature(state, x, arch)
    assert out.shape == (2, 60, 6, 25, 2) and feat.shape == (2, 256, 6, 25, 2)
    for mine, key in ((out, 'feat_out'), (feat, 'feat_feature')):
        ref = fix[key]
        assert np.abs(mg.probe(mine) - ref).max() <= 1e-5 * ref[0]


@pytest.mark.skipif(not refload.available(), reason='reference tree not mounted')
@pytest.mark.parametrize('name', ['ist_gcn', 'st_gcn'])
def test_oracle_matches_live_reference(name):
    g_args, num_class, shape = mg.MODEL_CASES[name]
    arch, state, x, label = _run_oracle(name)
    model = refload.build_reference_model(arch, shape[1], num_class, g_args, True)
    model.load_state_dict(state, strict=True)
    assert list(model.state_dict().keys()) == list(state.keys())
    model.train()
    ref = model(x)
    upd = {}
    mine = model_ref.forward(state, x, arch, training=True, update=upd)
    torch.testing.assert_close(mine, ref.detach(), rtol=1e-5, atol=2e-6)
    after = model.state_dict()
    for k, v in upd.items():
        torch.testing.assert_close(v, after[k], rtol=1e-5, atol=1e-6, msg=k)


def test_twostream_and_sgd_restatement():
    g = Graph('ntu-rgb+d', 'spatial')
    st = model_ref.make_state('st_gcn', 3, 60, g.A, seed=3)
    two = {('origin_stream.' + k): v for k, v in st.items()}
    two.update({('motion_stream.' + k): v for k, v in model_ref.make_state('st_gcn', 3, 60, g.A, seed=4).items()})
    x = torch.randn(1, 3, 16, 25, 2, generator=torch.Generator().manual_seed(0))
    m = model_ref.motion_stream_input(x)
    assert m.shape == x.shape and not m[:, :, 0].any() and not m[:, :, -1].any()
    torch.testing.assert_close(m[:, :, 5], x[:, :, 5] - 0.5 * x[:, :, 6] - 0.5 * x[:, :, 4])
    with torch.no_grad():
        y = model_ref.twostream_forward(two, x)
    assert y.shape == (1, 60)
    # SGD(momentum .9, nesterov, wd 1e-4) vs torch.optim on a toy parameter
    p = torch.nn.Parameter(torch.randn(5, generator=torch.Generator().manual_seed(1)))
    opt = torch.optim.SGD([p], lr=0.1, momentum=0.9, nesterov=True, weight_decay=1e-4)
    q, bufs = p.detach().clone(), [None]
    for step in range(3):
        grad = torch.full((5,), 0.3 * (step + 1))
        p.grad = grad.clone()
        opt.step()
        model_ref.sgd_nesterov_step([q], [grad], bufs, 0.1)
        torch.testing.assert_close(q, p.detach())
    assert model_ref.adjust_lr(0.1, [20, 40], 25) == pytest.approx(0.01)
