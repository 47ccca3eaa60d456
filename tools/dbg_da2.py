"""Numerics + timing probe of istgcn_gcn_tc_dvals (adjacency gradient of the graph convolution)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'ist-gcn_b200')); sys.path.insert(0, ROOT)
import torch
from istgcn._lib import call
from istgcn.sparse import SparsePattern
from net.utils.graph import Graph
dev = 'cuda'


def run(layout, strategy, Cin, Cout, frames, time_it=False):
    g = Graph(layout, strategy)
    A = torch.tensor(g.A + getattr(g, 'A2', 0 * g.A) + getattr(g, 'A3', 0 * g.A), dtype=torch.float32)
    K, V = A.shape[0], A.shape[1]
    pat = SparsePattern((A != 0).numpy(), dev)
    torch.manual_seed(3)
    x = torch.randn(frames * V, Cin, device=dev)
    dz = torch.randn(frames * V, Cout, device=dev)
    Wc = torch.randn(K * Cin, Cout, device=dev) * 0.05
    dvals = torch.zeros(pat.nnz, device=dev)

    def f():
        call('gcn_tc_dvals', dz, x, Wc, pat.dst_ptr, pat.dst_src, pat.dst_id, pat.nnz, dvals, frames, V, K, Cin, Cout)
    f(); torch.cuda.synchronize()
    G = torch.einsum('fwn,kcn->kfwc', dz.view(frames, V, Cout).double(), Wc.view(K, Cin, Cout).double())
    dA = torch.einsum('fvc,kfwc->kvw', x.view(frames, V, Cin).double(), G)
    ref = dA.reshape(-1)[pat.flat_idx]
    err = (dvals.double() - ref).abs().max().item() / ref.abs().max().item()
    msg = '%s/%s Cin %3d Cout %3d frames %6d: rel err %.2e' % (layout, strategy, Cin, Cout, frames, err)
    if time_it:
        for _ in range(3): f()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(10): f()
        e1.record(); torch.cuda.synchronize()
        msg += '  %.3f ms' % (e0.elapsed_time(e1) / 10)
    print(msg, flush=True)


if __name__ == '__main__':
    for fr in (1, 5, 7, 64, 601):
        run('ntu-rgb+d_sym', 'spatial_3_sym', 64, 64, fr)
    run('ntu-rgb+d_sym', 'spatial_3_sym', 64, 128, 333)
    run('ntu-rgb+d_sym', 'spatial_3_sym', 256, 128, 500)
    run('openpose_sym', 'spatial_3_sym', 64, 96, 500)
    run('ntu-rgb+d', 'spatial', 96, 64, 500)
    run('ntu-rgb+d', 'uniform', 64, 64, 100)
    frames = 128 * 300
    for Cin, Cout, fr in [(64, 64, frames), (64, 128, frames), (128, 128, frames // 2), (128, 256, frames // 2), (256, 256, frames // 4)]:
        run('ntu-rgb+d_sym', 'spatial_3_sym', Cin, Cout, fr, time_it=True)
