"""One training step of a golden case (for compute-sanitizer / poison checks).
usage: one_step.py <case> <math> [poison]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'ist-gcn_b200')); sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import torch, torch.nn.functional as F
import istgcn
import make_golden as mg
from net.utils.graph import Graph
import net.ist_gcn, net.st_gcn_mstcn_1x1
name, math = sys.argv[1], sys.argv[2]
poison = len(sys.argv) > 3
g_args, num_class, shape = mg.MODEL_CASES[name]
state = mg.case_state(name, Graph(**g_args))
x, label = mg.case_inputs(name, shape, num_class)
cls = net.ist_gcn.Model if name.startswith('ist_gcn') else net.st_gcn_mstcn_1x1.Model
dev = torch.device('cuda')
model = cls(shape[1], num_class, g_args, True)
model.load_state_dict(state); model = model.to(dev)
istgcn.set_math(math)
def step():
    model.zero_grad(set_to_none=True)
    logits = model(x.to(dev))
    F.cross_entropy(logits, label.to(dev)).backward()
    torch.cuda.synchronize()
    return {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}
def do_poison(val):
    junk = [torch.full((1 << 28,), val, device=dev)]
    for n in (128, 1024, 8192, 65536, 200000):
        junk += [torch.full((n,), val, device=dev) for _ in range(600)]
    del junk
if poison:
    do_poison(float('nan'))
g1 = step()
if poison:
    do_poison(1e30)
g2 = step()
bad = 0
for k in g1:
    for tag, g in (('nan-poison', g1[k]), ('1e30-poison', g2[k])):
        if not torch.isfinite(g).all() or g.abs().max().item() > 1e15:
            print('BAD', tag, k, g.abs().max().item()); bad += 1
print('done, bad =', bad)
