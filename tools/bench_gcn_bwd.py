#!/usr/bin/env python
"""Per-layer timing (CUDA-graph replays) of the graph-conv weight + adjacency gradient at the
BASELINE.json cfg-2 layer shapes: the one-pass pair kernel (csrc/gcn_pair_tc.cu) against the two
kernels it replaces (gcn_tc_dw2 + gcn_tc_da2 / gcn_tc_da).    python tools/bench_gcn_bwd.py [--iters 10]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, 'ist-gcn_b200'), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)
import torch  # noqa: E402

LAYERS = [(64, 64, 38400), (64, 128, 38400), (128, 128, 19200), (128, 256, 19200), (256, 256, 9600)]


def timed(fn, iters):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--iters', type=int, default=10)
    ap.add_argument('--only-pair', action='store_true')
    args = ap.parse_args()
    from istgcn import ops
    from istgcn._lib import call
    from istgcn.sparse import SparsePattern
    from net.utils.graph import Graph
    dev = torch.device('cuda')
    g = Graph('ntu-rgb+d_sym', 'spatial_3_sym')
    A = sum(torch.tensor(getattr(g, n), dtype=torch.float64) for n in ('A', 'A2', 'A3'))
    K, V = A.shape[0], A.shape[1]
    pat = SparsePattern((A != 0).numpy(), dev)
    vals = A.reshape(-1)[pat.flat_idx.cpu()].float().to(dev)
    for cin, cout, frames in LAYERS:
        x = torch.randn(frames * V, cin, device=dev)
        dz = torch.randn(frames * V, cout, device=dev)
        Wc = torch.randn(K * cin, cout, device=dev) * 0.05
        dW, dv = torch.zeros(K * cin, cout, device=dev), torch.zeros(pat.nnz, device=dev)
        items, ctas, joints = pat.pair_items(cin, cout)
        nb = int(items[0, 1] * items[0, 5])
        ws = torch.zeros(pat.npairs, cin, cout, device=dev)

        def pair():
            ws.zero_()
            call('gcn_pair_grads', dz, x, vals, Wc, items, items.shape[0], ctas, ctas.shape[0], joints, pat.pair_of,
                 pat.npairs, pat.entry_pair, pat.k_ptr, pat.nnz, ws, dW, dv, frames, V, K, cin, cout)

        def old():
            call('gcn_tc_dvals', dz, x, Wc, pat.dst_ptr, pat.dst_src, pat.dst_id, pat.nnz, dv, frames, V, K, cin, cout)
            call('gcn_tc_dw', dz, x, vals, pat.dst_ptr, pat.dst_src, pat.dst_id, pat.nnz, dW, None, frames, V, K,
                 cin, cout, 0, 0, 1, 0)

        t_pair = timed(pair, args.iters)
        t_old = float('nan') if args.only_pair else timed(old, args.iters)
        mb = 4 * frames * V * (cin + cout) / 1e6
        print('Cin=%3d Cout=%3d frames=%5d items=%3d nb=%3d  pair %7.1f us (%.0f GB/s on %.0f MB)   dw2+da %7.1f us' % (
            cin, cout, frames, items.shape[0], nb, t_pair, mb / t_pair * 1e3, mb, t_old))


if __name__ == '__main__':
    main()
