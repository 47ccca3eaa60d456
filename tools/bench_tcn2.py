#!/usr/bin/env python
"""Per-kernel timing of the streaming TCN kernels (csrc/tcn2.cu) at the BASELINE.json cfg-2 layer
shapes (batch 64 -> NM = 128, V = 25): CUDA events around each launch, inputs larger than L2,
algorithmic bytes / time against the measured HBM copy bandwidth.

    python tools/bench_tcn2.py [--iters 20] [--json out.json] [--only KERNEL]

Algorithmic bytes per launch (fp32): down r_in*C*4 (+h1), up r_out*C*4 (+h2), bwd_up 2*r_out*C*4
(+h2, dh2), bwd_down 2*r_in*C*4 (+dh1), conv / bwd_conv: the bp-wide tensors only.
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

LAYERS = [  # (C, b, T, stride) of the ten blocks of net.ist_gcn at T = 300
    (64, 8, 300, 1), (128, 11, 300, 2), (128, 11, 150, 1), (256, 16, 150, 2), (256, 16, 75, 1)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--iters', type=int, default=20)
    ap.add_argument('--json', default=None)
    ap.add_argument('--only', default=None)
    ap.add_argument('--nm', type=int, default=128)
    args = ap.parse_args()
    from istgcn import ops
    from istgcn._lib import call, i64, u64
    dev = torch.device('cuda')
    peak = 6553.3
    mp = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.isfile(mp):
        peak = float(json.load(open(mp)).get('hbm_gbs', peak))
    V, NM = 25, args.nm
    rows_out = []
    for C, b, T, s in LAYERS:
        bp = 8 if b <= 8 else 16
        Tout = (T - 1) // s + 1
        rin, rout = NM * T * V, NM * Tout * V
        g = torch.Generator(device='cuda').manual_seed(C + T)
        rnd = lambda *sh: torch.randn(*sh, device=dev, generator=g)          # noqa: E731
        z, go = rnd(rin, C), rnd(rout, C)
        co = [rnd(C) * 0.1 for _ in range(8)]
        Wd, bd = rnd(C, bp) * 0.1, rnd(bp) * 0.1
        Weff, beff = rnd(15, bp, bp) * 0.1, rnd(bp) * 0.1
        Wu, bu = rnd(bp, C) * 0.1, rnd(C) * 0.1
        h1, h2 = torch.empty(rin, bp, device=dev), torch.empty(rout, bp, device=dev)
        u, g1 = torch.empty(rout, C, device=dev), torch.empty(rin, C, device=dev)
        dh2, dh1 = torch.empty(rout, bp, device=dev), torch.empty(rin, bp, device=dev)
        st = torch.zeros(4, C, device=dev, dtype=torch.float64)
        dWu, dbu, dbe = torch.zeros(bp, C, device=dev), torch.zeros(C, device=dev), torch.zeros(bp, device=dev)
        dWe, dbd, dWd = torch.zeros(15, bp, bp, device=dev), torch.zeros(bp, device=dev), torch.zeros(C, bp, device=dev)
        sc = ops.step_counter(dev)
        kernels = [
            ('tcn2_down', lambda: call('tcn2_down', z, co[0], co[1], co[2], Wd, bd, h1, i64(rin), C, bp),
             4 * rin * (C + bp)),
            ('tcn2_conv', lambda: call('tcn2_conv', h1, Weff, beff, h2, NM, T, V, bp, s), 4 * (rin + rout) * bp),
            ('tcn2_up', lambda: call('tcn2_up', h2, Wu, bu, u, st[0], st[1], i64(rout), C, bp), 4 * rout * (C + bp)),
            ('tcn2_bwd_up', lambda: call('tcn2_bwd_up', go, u, co[3], co[4], co[5], co[6], h2, Wu, dh2, dWu, dbu, dbe,
                                         i64(rout), C, bp, 0.5, u64(77), sc), 4 * rout * (2 * C + 2 * bp)),
            ('tcn2_bwd_conv', lambda: call('tcn2_bwd_conv', dh2, h1, Weff, dh1, dWe, dbd, NM, T, V, bp, s),
             4 * (2 * rin + 2 * rout) * bp),
            ('tcn2_bwd_down', lambda: call('tcn2_bwd_down', dh1, z, co[0], co[1], co[2], co[7], Wd, g1, dWd, st[2],
                                           st[3], i64(rin), C, bp), 4 * rin * (2 * C + 2 * bp)),
        ]
        for name, fn, nbytes in kernels:
            if args.only and args.only != name:
                fn()                          # still produce the tensors the later kernels read
                continue
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            # the launches are replayed from a CUDA graph: the small kernels run for 10-30 us, less
            # than a Python -> ctypes -> launch round trip, so eager timing would measure the host
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                for _ in range(args.iters):
                    fn()
            graph.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            graph.replay()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.iters
            gbs = nbytes / ms / 1e6
            rows_out.append({'kernel': name, 'C': C, 'bp': bp, 'T': T, 'stride': s, 'us': ms * 1e3,
                             'algorithmic_MB': nbytes / 1e6, 'GBps': gbs, 'frac_of_hbm_peak': gbs / peak})
            print('%-14s C=%3d bp=%2d T=%3d s=%d  %8.1f us  %8.1f MB  %7.0f GB/s  %.2f of %.0f' % (
                name, C, bp, T, s, ms * 1e3, nbytes / 1e6, gbs, gbs / peak, peak))
        del z, go, u, g1
        torch.cuda.empty_cache()
    if args.json:
        with open(args.json, 'w') as f:
            json.dump({'hbm_peak_gbs': peak, 'nm': NM, 'rows': rows_out}, f, indent=1)


if __name__ == '__main__':
    main()
