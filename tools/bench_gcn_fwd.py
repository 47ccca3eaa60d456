#!/usr/bin/env python
"""Per-layer timing (CUDA-graph replays) and fp64 check of the fused graph convolution
`istgcn_gcn_tc` at the BASELINE.json cfg-2 layer shapes, forward form (bias term + BatchNorm sums)
and input-gradient form (transposed lists, reduce-add).  ISTGCN_GCN_TC_V2=1 selects the lane-mask
engine (csrc/gcn_tc2.cu) instead of the block-exchange engine (csrc/gcn_tc3.cu).

    python tools/bench_gcn_fwd.py [--iters 10] [--json out.json]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, 'ist-gcn_b200'), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)
import torch  # noqa: E402

LAYERS = [(64, 64, 38400), (64, 128, 19200), (128, 128, 19200), (128, 256, 9600), (256, 256, 9600)]


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


def rel(a, b):
    b = b.to(a.device).double()
    return ((a.double() - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--iters', type=int, default=10)
    ap.add_argument('--json', default=None)
    ap.add_argument('--check-frames', type=int, default=203)
    args = ap.parse_args()
    from istgcn._lib import call
    from istgcn.sparse import SparsePattern
    from net.utils.graph import Graph
    dev = torch.device('cuda')
    peak = 6553.3
    mp = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.isfile(mp):
        peak = float(json.load(open(mp)).get('hbm_gbs', peak))
    g = Graph('ntu-rgb+d_sym', 'spatial_3_sym')
    A = sum(torch.tensor(getattr(g, n), dtype=torch.float64) for n in ('A', 'A2', 'A3'))
    K, V = A.shape[0], A.shape[1]
    pat = SparsePattern((A != 0).numpy(), dev)
    vals = A.reshape(-1)[pat.flat_idx.cpu()].float().to(dev)
    colsum = A.sum(1).float().contiguous().to(dev)
    Ad = A.to(dev)
    rows = []
    for cin, cout, frames in LAYERS:
        x = torch.randn(frames * V, cin, device=dev)
        W2 = torch.randn(K * cout, cin, device=dev) * 0.05
        bias = torch.randn(K, cout, device=dev)
        z = torch.empty(frames * V, cout, device=dev)
        st = torch.zeros(2, cout, device=dev, dtype=torch.float64)

        def fwd(fr=frames):
            call('gcn_tc', x, None, None, None, None, None, W2, vals, pat.dst_ptr, pat.dst_src, pat.dst_id, pat.nnz,
                 bias, colsum, None, z, None, st[0], st[1], fr, V, K, cin, cin, cout, 0, 0, 1, 0, 0)

        # fp64 check on the first `check-frames` frames (ragged last tile: 203 = 4 * 50 + 3)
        fc = args.check_frames
        z.fill_(float('nan'))
        st.zero_()
        fwd(fc)
        xa = torch.einsum('fvc,kvw->kfwc', x[:fc * V].view(fc, V, cin).double(), Ad)
        ref = (torch.einsum('kfwc,knc->fwn', xa, W2.view(K, cout, cin).double()) +
               torch.einsum('kw,kn->wn', colsum.double(), bias.double())[None]).reshape(fc * V, cout)
        e_z, e_s, e_q = rel(z[:fc * V], ref), rel(st[0], ref.sum(0)), rel(st[1], (ref * ref).sum(0))
        untouched = bool(torch.isnan(z[fc * V:fc * V + 64]).all().item())
        # full size: every row finite and the sums consistent with the stored output
        st.zero_()
        fwd()
        torch.cuda.synchronize()
        e_full = rel(st[0], z.double().sum(0))
        t_f = timed(fwd, args.iters)

        dz, Wc = z, W2.view(K, cout, cin).permute(0, 2, 1).reshape(K * cin, cout).contiguous()
        gin = torch.zeros(frames * V, cin, device=dev)

        def bwd():
            call('gcn_tc', dz, None, None, None, None, None, Wc, vals, pat.t_ptr, pat.t_src, pat.t_id, pat.nnz, None,
                 None, gin, gin, None, None, None, frames, V, K, cout, cout, cin, 0, 0, 1, 0, 0)

        bwd()
        G = torch.einsum('fwn,kcn->kfwc', dz[:fc * V].view(fc, V, cout).double(), Wc.view(K, cin, cout).double())
        dx = torch.einsum('kfwc,kvw->fvc', G, Ad).reshape(fc * V, cin)
        e_b = rel(gin[:fc * V], dx)
        t_b = timed(bwd, args.iters)
        mb = 4 * frames * V * (cin + cout) / 1e6
        rows.append({'Cin': cin, 'Cout': cout, 'frames': frames, 'fwd_us': t_f, 'bwd_x_us': t_b, 'algorithmic_MB': mb,
                     'fwd_frac_of_hbm': mb / t_f * 1e3 / peak, 'bwd_x_frac_of_hbm': mb / t_b * 1e3 / peak,
                     'err_z': e_z, 'err_sum': e_s, 'err_sumsq': e_q, 'err_dx': e_b, 'err_sum_full': e_full,
                     'rows_beyond_untouched': untouched})
        print('Cin=%3d Cout=%3d frames=%5d  fwd %7.1f us (%.2f of HBM)  bwd-x %7.1f us (%.2f)  err z %.1e sum %.1e '
              'sq %.1e dx %.1e full-sum %.1e untouched %s' % (cin, cout, frames, t_f, mb / t_f * 1e3 / peak, t_b,
                                                             mb / t_b * 1e3 / peak, e_z, e_s, e_q, e_b, e_full, untouched))
    if args.json:
        with open(args.json, 'w') as f:
            json.dump({'hbm_peak_gbs': peak, 'engine': 'gcn_tc2' if os.environ.get('ISTGCN_GCN_TC_V2') else 'gcn_tc3',
                       'rows': rows}, f, indent=1)


if __name__ == '__main__':
    main()
