import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'ist-gcn_b200')); sys.path.insert(0, ROOT)
import torch, numpy as np
from istgcn._lib import call
from istgcn.sparse import SparsePattern
from net.utils.graph import Graph
torch.manual_seed(0)
dev = 'cuda'
g = Graph('ntu-rgb+d_sym', 'spatial_3_sym')
A = torch.tensor(g.A + g.A2 + g.A3, dtype=torch.float32)
K, V = 4, 25
pat = SparsePattern((A != 0).numpy(), dev)
vals = A.reshape(-1)[pat.flat_idx.cpu()].to(dev)
for (Cin, Cout, frames) in [(64, 64, 10), (64, 64, 5), (128, 128, 7), (256, 256, 5), (3, 64, 5)]:
    x = torch.randn(frames * V, Cin, device=dev)
    dz = torch.randn(frames * V, Cout, device=dev)
    dW = torch.zeros(K * Cin, Cout, device=dev); db = torch.zeros(V, Cout, device=dev)
    call('gcn_tc_dw', dz, x, vals, pat.dst_ptr, pat.dst_src, pat.dst_id, pat.nnz, dW, db, frames, V, K, Cin, Cout, 0, 0, 1, 0)
    torch.cuda.synchronize()
    xp = torch.einsum('fvc,kvw->fwkc', x.view(frames, V, Cin).double().cpu(), A.double())      # X'[f,w,k,ci]
    ref = torch.einsum('fwkc,fwo->kco', xp, dz.view(frames, V, Cout).double().cpu()).reshape(K * Cin, Cout)
    refb = dz.view(frames, V, Cout).double().cpu().sum(0)
    m = dW.double().cpu()
    print(Cin, Cout, frames, 'err', ((m - ref).abs().max() / ref.abs().max()).item(), 'bias err', ((db.double().cpu() - refb).abs().max() / refb.abs().max()).item(),
          '|mine|', m.abs().max().item(), '|ref|', ref.abs().max().item())
    if Cin == 64 and frames == 5:
        print('mine[:2,:6]', m[:2, :6].numpy().round(3)); print('ref [:2,:6]', ref[:2, :6].numpy().round(3))
        # does mine match a permutation? correlation of rows
        mt = m.reshape(K, Cin, Cout); rt = ref.reshape(K, Cin, Cout)
        print('nonzero frac mine', (m.abs() > 1e-6).double().mean().item())
