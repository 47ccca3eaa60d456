"""Scratch: one block, CUDA path vs fp64 oracle, next to fp32 oracle vs fp64 oracle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'ist-gcn_b200')); sys.path.insert(0, ROOT)
import torch
import istgcn
from net.ist_gcn import st_gcn
from net.utils.graph import Graph
from oracle import model_ref
cin, cout, stride, t, nm = [int(a) for a in sys.argv[1:6]]
mode = sys.argv[6] if len(sys.argv) > 6 else '3xtf32'
structured = len(sys.argv) > 7
residual = cin != 3
gen = torch.Generator().manual_seed(1)
g = Graph('ntu-rgb+d_sym', 'spatial_3_sym'); K, V = 4, 25
blk = st_gcn(cin, cout, (9, K), stride, residual=residual)
with torch.no_grad():
    for m in blk.modules():
        if isinstance(m, torch.nn.Conv2d):
            m.weight.normal_(0, 0.02, generator=gen); m.bias.normal_(0, 0.05, generator=gen)
        elif isinstance(m, torch.nn.BatchNorm2d):
            m.weight.normal_(1, 0.02, generator=gen); m.bias.normal_(0, 0.05, generator=gen)
state = {'b.' + k: v.detach().clone() for k, v in blk.state_dict().items()}
x = torch.randn(nm, cin, t, V, generator=gen).abs() if residual else torch.randn(nm, cin, t, V, generator=gen)
adjs = [torch.tensor(getattr(g, n), dtype=torch.float32) * (1 + 0.2 * torch.randn(K, V, V, generator=gen)) for n in ('A', 'A2', 'A3')]
m_imp = 1 + 0.3 * torch.randn(3, generator=gen)
tout = (t - 1) // stride + 1
if structured:   # gradient of a global average pool: constant over (t, v)
    gout = (torch.randn(nm, cout, 1, 1, generator=gen) / (tout * V)).expand(nm, cout, tout, V).contiguous()
else:
    gout = torch.randn(nm, cout, tout, V, generator=gen)

def oracle(dt):
    st = {k: (v.clone().to(dt).requires_grad_(True) if v.is_floating_point() and 'running' not in k else (v.to(dt) if v.is_floating_point() else v)) for k, v in state.items()}
    xx = x.clone().to(dt).requires_grad_(True); aa = [a.clone().to(dt).requires_grad_(True) for a in adjs]; mm = m_imp.clone().to(dt).requires_grad_(True)
    out = model_ref.block_forward(st, 'b.', 'ist_gcn', xx, aa, mm, (cin, cout, stride, residual), True, 0.0, None)
    out.backward(gout.to(dt))
    gr = {'x': xx.grad, 'm_imp': mm.grad}
    for i in range(3): gr['A%d' % i] = aa[i].grad
    for k, v in st.items():
        if getattr(v, 'grad', None) is not None: gr[k[2:]] = v.grad
    return out.detach().double(), {k: v.double() for k, v in gr.items()}
o64, g64 = oracle(torch.float64); o32, g32 = oracle(torch.float32)
blk = blk.cuda().train(); istgcn.set_math(mode)
xg = x.cuda().requires_grad_(True); ag = [a.cuda().requires_grad_(True) for a in adjs]; mg = m_imp.cuda().requires_grad_(True)
out = blk(xg, ag[0], ag[1], ag[2], mg)[0]; out.backward(gout.cuda())
gm = {'x': xg.grad, 'm_imp': mg.grad}
for i in range(3): gm['A%d' % i] = ag[i].grad
for k, p in blk.named_parameters():
    if p.grad is not None: gm[k] = p.grad
union = (adjs[0] != 0) | (adjs[1] != 0) | (adjs[2] != 0)
def l2(a, b): return ((a - b).norm() / b.norm().clamp_min(1e-300)).item()
print('out: mine %.2e fp32 %.2e' % (l2(out.detach().cpu().double(), o64), l2(o32, o64)))
gmax = max(v.abs().max().item() for v in g64.values())
for k in sorted(g64):
    a, b, c = gm[k].detach().cpu().double(), g64[k], g32[k]
    if k.startswith('A'): a, b, c = a * union, b * union, c * union
    if b.abs().max().item() < 1e-9 * gmax:
        print('%-28s zero-grad: mine max %.2e fp32 max %.2e' % (k, a.abs().max().item(), c.abs().max().item())); continue
    print('%-28s mine %.2e  fp32 %.2e   |g| %.2e' % (k, l2(a, b), l2(c, b), b.norm().item()))
