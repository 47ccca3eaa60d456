import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'ist-gcn_b200')); sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import importlib, numpy as np, torch, torch.nn.functional as F
import istgcn
import make_golden as mg
from net.utils.graph import Graph
name, math = sys.argv[1], sys.argv[2]
g_args, num_class, shape = mg.MODEL_CASES[name]
state = mg.case_state(name, Graph(**g_args))
x, label = mg.case_inputs(name, shape, num_class)
dev = torch.device('cuda')
istgcn.set_math(math)
res = {}
modes = ('off', 'on') if len(sys.argv) < 4 else ('off', 'pert')
for mode in modes:
    if mode != 'on': os.environ['ISTGCN_GCN_SMALL_OFF'] = '1'
    else: os.environ.pop('ISTGCN_GCN_SMALL_OFF', None)
    if mode == 'pert': x = x * (1 + 1e-7 * torch.randn(x.shape))
    model = importlib.import_module('net.' + name).Model(shape[1], num_class, g_args, True)
    model.load_state_dict(state); model = model.to(dev).train()
    outs = []
    gouts = {}
    for b in model.st_gcn_networks:
        def wrap(f):
            def g(*a, **k):
                o = f(*a, **k); outs.append(o.detach().clone()); o.register_hook(lambda g, i=len(outs) - 1: gouts.__setitem__(i, g.detach().clone())); return o
            return g
        b.forward_cl = wrap(b.forward_cl)
    logits = model(x.to(dev))
    F.cross_entropy(logits, label.to(dev)).backward()
    res[mode + '_g'] = gouts
    res[mode] = (outs, {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}, logits.detach())
o0, g0, l0 = res['off']; o1, g1, l1 = res[modes[1]]
for i, (a, b) in enumerate(zip(o0, o1)):
    d = (a - b).abs()
    flips = ((a > 0) != (b > 0)).sum().item()
    print('block %d out: max abs diff %.3e (max %.3e) relu-support flips %d of %d' % (i, d.max().item(), a.abs().max().item(), flips, a.numel()))
for i in sorted(res['off_g']):
    a, b = res['off_g'][i], res[modes[1] + '_g'][i]
    d = (a - b).abs()
    print('grad wrt block %d out: rel l2 %.3e  max abs %.3e (max %.3e) nonzero-diff elems %d' % (i, (a - b).norm().item() / a.norm().item(), d.max().item(), a.abs().max().item(), (d > 1e-3 * a.abs().max()).sum().item()))
print('logits diff %.3e' % (l0 - l1).abs().max().item())
for k in g0:
    e = (g0[k] - g1[k]).norm().item() / max(g0[k].norm().item(), 1e-30)
    if e > 1e-4: print('grad %-45s rel l2 diff %.3e' % (k, e))
