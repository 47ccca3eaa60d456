"""Numerics + timing probe of the tcgen05 graph-convolution entry point (istgcn_gcn_tc) against a
plain fp32 torch evaluation of tgcn.py:76-89 on the same device."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'ist-gcn_b200')); sys.path.insert(0, ROOT)
import torch
from istgcn._lib import call
from istgcn.sparse import SparsePattern
from net.utils.graph import Graph

dev = 'cuda'
torch.backends.cuda.matmul.allow_tf32 = False


def run(layout, strategy, Cin, Cout, frames, reduce=False, time_it=False):
    g = Graph(layout, strategy)
    A = torch.tensor(g.A + getattr(g, 'A2', 0 * g.A) + getattr(g, 'A3', 0 * g.A), dtype=torch.float32)
    K, V = A.shape[0], A.shape[1]
    pat = SparsePattern((A != 0).numpy(), dev)
    vals = A.reshape(-1)[pat.flat_idx.cpu()].to(dev)
    torch.manual_seed(1)
    x = torch.randn(frames * V, Cin, device=dev)
    W2 = torch.randn(K * Cout, Cin, device=dev) * 0.05
    bias = torch.randn(K, Cout, device=dev)
    colsum = A.sum(1).contiguous().to(dev)
    z0 = torch.randn(frames * V, Cout, device=dev)
    z = z0.clone() if reduce else torch.full((frames * V, Cout), float('nan'), device=dev)
    st = torch.zeros(2, Cout, device=dev, dtype=torch.float64)

    def f():
        call('gcn_tc', x, None, None, None, None, None, W2, vals, pat.dst_ptr, pat.dst_src, pat.dst_id,
             pat.nnz, None if reduce else bias, None if reduce else colsum, z if reduce else None, z, None,
             None if reduce else st[0], None if reduce else st[1], frames, V, K, Cin, Cin, Cout, 0, 0, 1, 0, 0)
    f()
    torch.cuda.synchronize()
    Ad = A.to(dev)
    xa = torch.einsum('fvc,kvw->kfwc', x.view(frames, V, Cin).double(), Ad.double())      # aggregated
    ref = torch.einsum('kfwc,knc->fwn', xa, W2.view(K, Cout, Cin).double())
    if not reduce:
        ref = ref + torch.einsum('kw,kn->wn', colsum.double(), bias.double())[None]
    else:
        ref = ref + z0.view(frames, V, Cout).double()
    ref = ref.reshape(frames * V, Cout)
    err = (z.double() - ref).abs().max().item() / ref.abs().max().item()
    msg = '%s/%s Cin %3d Cout %3d frames %6d reduce %d: rel err %.2e' % (layout, strategy, Cin, Cout, frames, reduce, err)
    if not reduce:
        e1 = (st[0] - ref.sum(0)).abs().max().item() / ref.sum(0).abs().max().item()
        e2 = (st[1] - (ref * ref).sum(0)).abs().max().item() / (ref * ref).sum(0).abs().max().item()
        msg += ' stats %.1e %.1e' % (e1, e2)
    if time_it:
        for _ in range(3): f()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(10): f()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        gbs = frames * V * (Cin + Cout) * 4 / ms / 1e6
        msg += '  %.3f ms  %.0f GB/s' % (ms, gbs)
    print(msg, flush=True)
    return err


if __name__ == '__main__':
    worst = 0.0
    quick = os.environ.get('QUICK')
    for fr in (() if quick else (1, 3, 4, 7, 64)):
        worst = max(worst, run('ntu-rgb+d_sym', 'spatial_3_sym', 64, 64, fr))
    if not quick:
        worst = max(worst, run('ntu-rgb+d_sym', 'spatial_3_sym', 64, 128, 601))
        worst = max(worst, run('ntu-rgb+d_sym', 'spatial_3_sym', 128, 256, 333))
        worst = max(worst, run('ntu-rgb+d_sym', 'spatial_3_sym', 256, 256, 2001))
        worst = max(worst, run('ntu-rgb+d_sym', 'spatial_3_sym', 128, 64, 500, reduce=True))
        worst = max(worst, run('openpose_sym', 'spatial_3_sym', 64, 96, 500))
        worst = max(worst, run('ntu-rgb+d', 'spatial', 64, 64, 500))
    print('worst rel err %.2e' % worst)
    frames = 128 * 300
    for Cin, Cout, fr in [(64, 64, frames), (128, 128, frames // 2), (256, 256, frames // 4),
                          (64, 128, frames), (128, 256, frames // 2)]:
        run('ntu-rgb+d_sym', 'spatial_3_sym', Cin, Cout, fr, time_it=True)
