import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'ist-gcn_b200')); sys.path.insert(0, ROOT)
import torch
from istgcn._lib import call
from istgcn import _lib
from istgcn.sparse import SparsePattern
from net.utils.graph import Graph
dev = 'cuda'
g = Graph('ntu-rgb+d_sym', 'spatial_3_sym')
A = torch.tensor(g.A + g.A2 + g.A3, dtype=torch.float32)
K, V = 4, 25
pat = SparsePattern((A != 0).numpy(), dev)
names = {0: 'tma.wait_empty', 1: 'tma.other', 2: 'mma.wait_t_empty', 3: 'mma.wait_full', 4: 'mma.other',
         5: 'epi.wait_t_full', 6: 'epi.wait_x_full', 7: 'epi.work', 8: 'ld.wait_x_empty', 9: 'ld.work'}
for C, frames in [(64, 38400), (128, 19200), (256, 9600)]:
    x = torch.randn(frames * V, C, device=dev); dz = torch.randn(frames * V, C, device=dev)
    Wc = torch.randn(K * C, C, device=dev) * 0.05; dv = torch.zeros(pat.nnz, device=dev)
    f = lambda: call('gcn_tc_dvals', dz, x, Wc, pat.dst_ptr, pat.dst_src, pat.dst_id, pat.nnz, dv, frames, V, K, C, C)
    for _ in range(3): f()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(10): f()
    e1.record(); torch.cuda.synchronize()
    h = _lib.load(); out = (ctypes.c_ulonglong * 32)()
    if hasattr(h, 'istgcn_debug_prof_da'):
        h.istgcn_debug_prof_da(out); f(); h.istgcn_debug_prof_da(out)
        print('   ' + '  '.join('%s=%dk' % (names[i], out[i] // 1000) for i in sorted(names)))
    print('C %3d: %.3f ms' % (C, e0.elapsed_time(e1) / 10))
