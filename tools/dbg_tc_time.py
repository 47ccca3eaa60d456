import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'ist-gcn_b200')); sys.path.insert(0, ROOT)
import torch, numpy as np
from istgcn._lib import call
from istgcn import _lib
import ctypes
from istgcn.sparse import SparsePattern
from net.utils.graph import Graph
dev = 'cuda'
g = Graph('ntu-rgb+d_sym', 'spatial_3_sym')
Afull = torch.tensor(g.A + g.A2 + g.A3, dtype=torch.float32)
K, V = 4, 25
def run(A, Cin, Cout, frames, tag):
    mask = (A != 0).numpy()
    pat = SparsePattern(mask, dev)
    vals = A.reshape(-1)[pat.flat_idx.cpu()].to(dev) if pat.nnz else torch.zeros(1, device=dev)
    x = torch.randn(frames * V, Cin, device=dev); W2 = torch.randn(K * Cout, Cin, device=dev) * 0.05
    bias = torch.randn(K, Cout, device=dev); colsum = A.sum(1).contiguous().to(dev); z = torch.empty(frames * V, Cout, device=dev)
    st = torch.zeros(2, Cout, device=dev, dtype=torch.float64)
    def f():
        call('gcn_tc', x, None, None, None, None, None, W2, vals, pat.dst_ptr, pat.dst_src, pat.dst_id, pat.nnz, bias, colsum, None, z, None, st[0], st[1], frames, V, K, Cin, Cin, Cout, 0, 0, 1, 0, 0)
    for _ in range(3): f()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(10): f()
    e1.record(); torch.cuda.synchronize()
    h = _lib.load(); out = (ctypes.c_ulonglong * 32)()
    if not hasattr(h, 'istgcn_debug_prof'):
        print('%-10s Cin %3d Cout %3d nnz %3d: %.3f ms' % (tag, Cin, Cout, pat.nnz, e0.elapsed_time(e1) / 10)); return
    h.istgcn_debug_prof(out); f(); h.istgcn_debug_prof(out)
    names = {0:'tma.wait_b_empty',1:'tma.other',2:'mma.wait_t_empty',3:'mma.wait_a_full',4:'mma.wait_b_full',5:'mma.other',7:'epi.wait_t_full',8:'epi.work',9:'ld.wait_x_empty',10:'ld.work',12:'agg.wait_x_full',13:'agg.wait_a_empty',14:'agg.work',15:'agg.fence_arrive'}
    print('   ' + '  '.join('%s=%dk' % (names[i], out[i] // 1000) for i in sorted(names)))
    print('%-10s Cin %3d Cout %3d nnz %3d: %.3f ms' % (tag, Cin, Cout, pat.nnz, e0.elapsed_time(e1) / 10))
frames = 128 * 300
for Cin, Cout, fr in [(64, 64, frames), (128, 128, frames // 2), (256, 256, frames // 4)]:
    run(Afull, Cin, Cout, fr, 'full')
