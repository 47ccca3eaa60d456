"""Scratch: error of d(loss)/d(block output i) along the backward chain: CUDA path and fp32
oracle, both vs the fp64 oracle."""
import importlib.util, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'ist-gcn_b200')); sys.path.insert(0, ROOT)
import torch, torch.nn.functional as F
import istgcn, net.ist_gcn
from istgcn import modules
from oracle import model_ref
from net.utils.graph import Graph
spec = importlib.util.spec_from_file_location('mg', os.path.join(ROOT, 'tests/golden/make_golden.py'))
mg = importlib.util.module_from_spec(spec); spec.loader.exec_module(mg)
name = 'ist_gcn'; mode = sys.argv[1] if len(sys.argv) > 1 else '3xtf32'
g_args, ncls, shape = mg.MODEL_CASES[name]
state = mg.case_state(name, Graph(**g_args)); x, label = mg.case_inputs(name, shape, ncls)

def oracle(dt):
    st = {k: (v.clone().to(dt).requires_grad_(True) if v.is_floating_point() and 'running' not in k and k not in ('A', 'A2', 'A3')
              else (v.to(dt) if v.is_floating_point() else v)) for k, v in state.items()}
    xx = x.to(dt)
    N, C, T, V, M = xx.shape
    h = xx.permute(0, 4, 3, 1, 2).contiguous().view(N * M, V * C, T)
    h = model_ref._bn(st, 'data_bn.', h, True)
    h = h.view(N, M, V, C, T).permute(0, 1, 3, 4, 2).contiguous().view(N * M, C, T, V)
    outs = []
    for i, cfg in enumerate(model_ref.block_table('ist_gcn', C)):
        adjs = [st['A'] * st['edge_importance.%d' % i], st['A2'] * st['edge_importance2.%d' % i], st['A3'] * st['edge_importance3.%d' % i]]
        h = model_ref.block_forward(st, 'st_gcn_networks.%d.' % i, 'ist_gcn', h, adjs, st['mstcn_importance.%d' % i], cfg, True)
        h.retain_grad(); outs.append(h)
    y = F.avg_pool2d(h, h.shape[2:]).view(N, M, -1, 1, 1).mean(dim=1)
    y = F.conv2d(y, st['fcn.weight'], st['fcn.bias']).view(N, -1)
    F.cross_entropy(y, label).backward()
    return [o.grad.double().permute(0, 2, 3, 1) for o in outs], [o.detach().double().permute(0, 2, 3, 1) for o in outs]

g64, a64 = oracle(torch.float64); g32, a32 = oracle(torch.float32)
m = net.ist_gcn.Model(shape[1], ncls, g_args, True); m.load_state_dict(state); m = m.cuda().train()
istgcn.set_math(mode)
grads, acts = {}, {}
orig = modules.FusedBlockMixin.forward_cl
def patched(self, xx, adjs, m_imp, pattern):
    out = orig(self, xx, adjs, m_imp, pattern)
    idx = list(m.st_gcn_networks).index(self)
    acts[idx] = out.detach().cpu().double()
    out.register_hook(lambda g, idx=idx: grads.__setitem__(idx, g.detach().cpu().double()))
    return out
modules.FusedBlockMixin.forward_cl = patched
F.cross_entropy(m(x.cuda()), label.cuda()).backward()
def l2(a, b): return ((a - b).norm() / b.norm().clamp_min(1e-300)).item()
print('blk   act mine   act fp32 | grad mine  grad fp32')
for i in range(9, -1, -1):
    print('%d    %.2e   %.2e | %.2e   %.2e' % (i, l2(acts[i], a64[i]), l2(a32[i], a64[i]), l2(grads[i], g64[i]), l2(g32[i], g64[i])))

# ---- single block with the real data of block `bi`
bi = int(sys.argv[2]) if len(sys.argv) > 2 else 9
cfgs = model_ref.block_table('ist_gcn', 3)
cin, cout, stride, residual = cfgs[bi]
xin = acts[bi - 1].float()                       # (NM, T, V, C) channels-last, the CUDA path's own
gout = g64[bi].float()
pfx = 'st_gcn_networks.%d.' % bi
def single(dt):
    st = {k: (v.clone().to(dt).requires_grad_(True) if v.is_floating_point() and 'running' not in k and k not in ('A', 'A2', 'A3')
              else (v.to(dt) if v.is_floating_point() else v)) for k, v in state.items() if k.startswith(pfx) or 'importance' in k or k in ('A', 'A2', 'A3')}
    xx = xin.permute(0, 3, 1, 2).contiguous().to(dt).requires_grad_(True)
    adjs = [st['A'] * st['edge_importance.%d' % bi], st['A2'] * st['edge_importance2.%d' % bi], st['A3'] * st['edge_importance3.%d' % bi]]
    out = model_ref.block_forward(st, pfx, 'ist_gcn', xx, adjs, st['mstcn_importance.%d' % bi], cfgs[bi], True)
    out.backward(gout.permute(0, 3, 1, 2).contiguous().to(dt))
    gr = {k: v.grad.double() for k, v in st.items() if getattr(v, 'grad', None) is not None}
    gr['x'] = xx.grad.double().permute(0, 2, 3, 1)
    return gr
s64, s32 = single(torch.float64), single(torch.float32)
blk = m.st_gcn_networks[bi]
for p_ in m.parameters(): p_.grad = None
modules.FusedBlockMixin.forward_cl = orig
xg = xin.cuda().requires_grad_(True)
adjs = [m.A * m.edge_importance[bi], m.A2 * m.edge_importance2[bi], m.A3 * m.edge_importance3[bi]]
out = blk.forward_cl(xg, adjs, m.mstcn_importance[bi], m._pattern())
out.backward(gout.cuda())
mine = {k: p_.grad.detach().cpu().double() for k, p_ in m.named_parameters() if p_.grad is not None}
mine['x'] = xg.grad.detach().cpu().double()
gmax = max(v.abs().max().item() for v in s64.values())
print('single block %d: param, mine L2, fp32 L2, |g|' % bi)
for k in sorted(s64):
    if s64[k].abs().max().item() < 1e-7 * gmax: continue
    print('%-44s %.2e  %.2e  %.2e' % (k, l2(mine[k], s64[k]), l2(s32[k], s64[k]), s64[k].norm().item()))
