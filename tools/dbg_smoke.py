import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'ist-gcn_b200')); sys.path.insert(0, ROOT)
import torch, torch.nn.functional as F
import istgcn
from net.ist_gcn import Model
from oracle import model_ref
torch.manual_seed(0)
dev = torch.device('cuda:0')
g_args = dict(layout='ntu-rgb+d_sym', strategy='spatial_3_sym')
model = Model(3, 60, g_args, True)
state = model_ref.perturb_state(model_ref.make_state('ist_gcn', 3, 60, model.graph.A, model.graph.A2, model.graph.A3, seed=5))
model.load_state_dict(state, strict=True)
model = model.to(dev).train()
x = torch.randn(2, 3, 20, 25, 2)
label = torch.randint(0, 60, (2,))
istgcn.set_math(sys.argv[1] if len(sys.argv) > 1 else '3xtf32')
leaves = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and 'running' not in k and k not in ('A', 'A2', 'A3') else v) for k, v in state.items()}
ref = model_ref.forward(leaves, x, 'ist_gcn', training=True)
F.cross_entropy(ref, label).backward()
l64 = {k: (v.double().clone().requires_grad_(True) if v.is_floating_point() and 'running' not in k and k not in ('A', 'A2', 'A3') else (v.double() if v.is_floating_point() else v)) for k, v in state.items()}
r64 = model_ref.forward(l64, x.double(), 'ist_gcn', training=True)
F.cross_entropy(r64, label).backward()
for rep in range(3):
    model.zero_grad(set_to_none=True)
    logits = model(x.to(dev))
    loss = F.cross_entropy(logits, label.to(dev))
    loss.backward()
    torch.cuda.synchronize()
    err = (logits.detach().cpu() - ref.detach()).abs().max().item() / ref.detach().abs().max().item()
    bad = []
    for name, p in model.named_parameters():
        if leaves[name].grad is None: continue
        r = leaves[name].grad
        e = (p.grad.cpu() - r).abs().max().item() / max(r.abs().max().item(), 1e-12)
        if e > 2e-3: bad.append((name, e))
    dot = n1 = n2 = 0.0; worst = 0.0; w32 = 0.0
    for name, p in model.named_parameters():
        if l64[name].grad is None: continue
        a, b = p.grad.cpu().double(), l64[name].grad
        dot += (a * b).sum().item(); n1 += (a * a).sum().item(); n2 += (b * b).sum().item()
        worst = max(worst, ((a - b).norm() / b.norm().clamp_min(1e-30)).item() if b.abs().max() > 1e-9 else 0.0)
        c = leaves[name].grad.double(); w32 = max(w32, ((c - b).norm() / b.norm().clamp_min(1e-30)).item() if b.abs().max() > 1e-9 else 0.0)
    print('vs fp64: cosine %.8f worst rel-L2 %.3e (fp32 oracle vs fp64 worst rel-L2 %.3e)' % (dot / (n1 * n2) ** 0.5, worst, w32))
    print('rep', rep, 'logits err %.2e' % err, 'bad grads:', bad[:8], len(bad))
