"""Repeat one training step of a golden case; report, per repetition, how the block outputs and
the gradients differ from repetition 0 (hunting run-to-run nondeterminism)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'ist-gcn_b200')); sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import numpy as np, torch, torch.nn.functional as F
import istgcn
import make_golden as mg
from net.utils.graph import Graph
import net.ist_gcn
name, math = sys.argv[1], sys.argv[2]
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 8
g_args, num_class, shape = mg.MODEL_CASES[name]
state = mg.case_state(name, Graph(**g_args))
x, label = mg.case_inputs(name, shape, num_class)
dev = torch.device('cuda')
istgcn.set_math(math)
base = None
for rep in range(reps):
    model = net.ist_gcn.Model(shape[1], num_class, g_args, True)
    model.load_state_dict(state); model = model.to(dev)
    outs = []
    for b in model.st_gcn_networks:
        def wrap(f):
            def g(*a, **k):
                o = f(*a, **k); outs.append(o.detach().clone()); return o
            return g
        b.forward_cl = wrap(b.forward_cl)
    model.eval()
    with torch.no_grad():
        model(x.to(dev))
    outs.clear()
    model.train()
    logits = model(x.to(dev))
    loss = F.cross_entropy(logits, label.to(dev)); loss.backward()
    grads = {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}
    if base is None:
        base = (outs, grads); continue
    fd = ['%.1e/%d' % ((a - b).abs().max().item(), int(((a > 0) != (b > 0)).sum())) for a, b in zip(outs, base[0])]
    gmax = max(v.abs().max().item() for v in base[1].values())
    gd = sorted(((( grads[k] - base[1][k]).norm() / base[1][k].norm().clamp_min(1e-3 * gmax * base[1][k].numel() ** 0.5)).item(), k) for k in grads)[-3:]
    print(rep, 'loss %.9f' % loss.item(), 'fwd maxdiff/flips:', ' '.join(fd))
    print('    grad rel-L2 top:', '; '.join('%s=%.2e' % (k.replace('st_gcn_networks', 'blk'), v) for v, k in gd))
