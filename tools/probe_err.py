"""Top gradient-probe errors of a golden case vs the committed fixture.  usage: probe_err.py <case> <math> [reps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'ist-gcn_b200')); sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import numpy as np, torch, torch.nn.functional as F
import istgcn
import make_golden as mg
from net.utils.graph import Graph
import net.ist_gcn, net.st_gcn_mstcn_1x1
name, math = sys.argv[1], sys.argv[2]
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
g_args, num_class, shape = mg.MODEL_CASES[name]
state = mg.case_state(name, Graph(**g_args))
x, label = mg.case_inputs(name, shape, num_class)
fix = np.load(os.path.join(ROOT, 'tests', 'golden', 'model_%s.npz' % name))
cls = net.ist_gcn.Model if name.startswith('ist_gcn') else net.st_gcn_mstcn_1x1.Model
dev = torch.device('cuda')
istgcn.set_math(math)
names = [str(s) for s in fix['grad_names']]
gmax = max(np.abs(fix['grad|' + k][2:]).max() for k in names)
for rep in range(reps):
    model = cls(shape[1], num_class, g_args, True)
    model.load_state_dict(state); model = model.to(dev)
    model.eval()
    with torch.no_grad():
        model(x.to(dev))
    model.train()
    logits = model(x.to(dev))
    F.cross_entropy(logits, label.to(dev)).backward()
    params = dict(model.named_parameters())
    errs = {}
    for k in names:
        ref = fix['grad|' + k]
        mine = mg.probe(params[k].grad.cpu())
        errs[k] = np.abs(mine[2:] - ref[2:]).max() / max(np.abs(ref[2:]).max(), 1e-2 * gmax)
    top = sorted(((v, k) for k, v in errs.items()), reverse=True)[:4]
    print(rep, 'loss %.9f' % float(F.cross_entropy(logits, label.to(dev))), '; '.join('%s=%.3f' % (k.replace('st_gcn_networks', 'blk'), v) for v, k in top))
