"""Scratch diagnostic: gradient error of the CUDA path vs the fp64 oracle, next to the error of
plain fp32 PyTorch (CPU and GPU eager) vs the same fp64 oracle."""
import importlib.util, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'ist-gcn_b200')); sys.path.insert(0, ROOT)
import torch, torch.nn.functional as F
import istgcn, net.ist_gcn, net.st_gcn_mstcn_1x1
from oracle import model_ref
from net.utils.graph import Graph
spec = importlib.util.spec_from_file_location('mg', os.path.join(ROOT, 'tests/golden/make_golden.py'))
mg = importlib.util.module_from_spec(spec); spec.loader.exec_module(mg)
name = sys.argv[1] if len(sys.argv) > 1 else 'ist_gcn'
mode = sys.argv[2] if len(sys.argv) > 2 else '3xtf32'
g_args, ncls, shape = mg.MODEL_CASES[name]
state = mg.case_state(name, Graph(**g_args)); x, label = mg.case_inputs(name, shape, ncls)
if len(sys.argv) > 4:
    shape = (int(sys.argv[3]), shape[1], int(sys.argv[4]), shape[3], shape[4])
    gen = torch.Generator().manual_seed(5)
    x = torch.randn(shape, generator=gen); label = torch.randint(0, ncls, (shape[0],), generator=gen)
torch.backends.cuda.matmul.allow_tf32 = False; torch.backends.cudnn.allow_tf32 = False

def oracle(dtype, dev):
    lv = {k: (v.detach().clone().to(dev, dtype).requires_grad_(True) if v.is_floating_point() and 'running' not in k and k not in ('A','A2','A3')
              else (v.to(dev, dtype) if v.is_floating_point() else v.to(dev))) for k, v in state.items()}
    out = model_ref.forward(lv, x.to(dev, dtype), name.replace('_kinetics', ''), training=True)
    F.cross_entropy(out, label.to(dev)).backward()
    return out.detach().cpu().double(), {k: v.grad.detach().cpu().double() for k, v in lv.items() if getattr(v, 'grad', None) is not None}

o64, g64 = oracle(torch.float64, 'cpu')
o32, g32 = oracle(torch.float32, 'cpu')
o32g, g32g = oracle(torch.float32, 'cuda')
m = (net.ist_gcn if 'ist_gcn' in name else net.st_gcn_mstcn_1x1).Model(shape[1], ncls, g_args, True); m.load_state_dict(state); m = m.cuda().train()
istgcn.set_math(mode)
out = m(x.cuda()); F.cross_entropy(out, label.cuda()).backward()
gm = {k: p.grad.detach().cpu().double() for k, p in m.named_parameters() if p.grad is not None}
def l2(a, b): return ((a - b).norm() / b.norm().clamp_min(1e-300)).item()
def mx(a, b): return ((a - b).abs().max() / b.abs().max().clamp_min(1e-300)).item()
print('logits: mine %.2e cpu32 %.2e gpu32 %.2e' % (mx(out.detach().cpu().double(), o64), mx(o32, o64), mx(o32g, o64)))
rows = []
gmax = max(v.abs().max().item() for v in g64.values())
for k in g64:
    if g64[k].abs().max().item() < 1e-6 * gmax: continue
    rows.append((l2(gm[k], g64[k]), l2(g32[k], g64[k]), l2(g32g[k], g64[k]), mx(gm[k], g64[k]), g64[k].norm().item(), k))
rows.sort(reverse=True)
print('%-48s %9s %9s %9s %9s %9s' % ('param', 'mine L2', 'cpu32 L2', 'gpu32 L2', 'mine max', '|g|'))
for r in rows[:25]:
    print('%-48s %9.2e %9.2e %9.2e %9.2e %9.2e' % (r[5], r[0], r[1], r[2], r[3], r[4]))
import statistics
print('median mine L2 %.2e cpu32 %.2e gpu32 %.2e' % tuple(statistics.median(r[i] for r in rows) for i in range(3)))
