"""torch.autograd Functions over the C ABI (include/istgcn_b200.h).

Layout: activations are channels-last fp32 tensors of shape (N*M, T, V, C) -- the reference's
(N*M, C, T, V) (net/st_gcnold.py:80) with C moved to the fastest axis -- so a 1x1 convolution is
a row-major GEMM and a temporal tap is a shift of V rows.

One Function per st_gcn block (``STBlock``): forward launches the fused graph convolution
(+ strided 1x1 residual conv), the Inception-TCN kernels and the block tail, with the
training-mode BatchNorm statistics accumulated in the producers' epilogues; backward launches
the mirrored chain.  Parameter regrouping (A*importance, merged temporal taps, weight
transposes) is done with tiny differentiable torch ops *outside* the Function, so autograd
maps the kernel's gradients back onto the reference's parameter tensors.
"""
import os

import torch
from torch.autograd import Function

from . import _lib, math_flag, use_tc
from ._lib import call, f64, i64, u64


_step_counters = {}


def step_counter(device):
    """Device-resident training-step counter (int64[1]) mixed into every dropout seed; the
    trainer increments it once per step (inside the captured graph when graphs are used)."""
    device = torch.device(device)
    index = device.index if device.index is not None else torch.cuda.current_device()
    key = (device.type, index)
    device = torch.device(device.type, index)
    if key not in _step_counters:
        _step_counters[key] = torch.zeros(1, dtype=torch.int64, device=device)
    return _step_counters[key]


def _coeffs(n, C, device):
    buf = torch.empty(n, C, device=device, dtype=torch.float32)
    return [buf[i] for i in range(n)]


class BNState(object):
    """The mutable bits of an nn.BatchNorm module that the kernels touch."""
    __slots__ = ('running_mean', 'running_var', 'momentum', 'eps')

    def __init__(self, bn):
        self.running_mean, self.running_var = bn.running_mean, bn.running_var
        self.momentum = 0.1 if bn.momentum is None else float(bn.momentum)
        self.eps = float(bn.eps)


def _bn_forward_coeffs(training, sum_, sumsq, count, weight, st, C, device):
    """-> mean, scale, rstd for y = (x - mean)*scale + beta (rstd is None in eval mode)."""
    if training:
        scale, mean, rstd = _coeffs(3, C, device)
        call('bn_finalize', sum_, sumsq, f64(count), weight, st.running_mean, st.running_var,
             st.momentum, st.eps, scale, mean, rstd, C)
        return mean, scale, rstd
    scale, = _coeffs(1, C, device)
    call('bn_eval_coeffs', weight, st.running_var, st.eps, scale, C)
    return st.running_mean, scale, None


def _bn_fold_enabled():
    """BatchNorm bookkeeping folded into the consumer kernels (istgcn_*_bn entry points) instead of the
    one-block bn_finalize / bn_bwd_coeffs launches; ISTGCN_BN_FOLD=0 keeps the separate launches."""
    return os.environ.get('ISTGCN_BN_FOLD', '1') != '0'


def _small_bwd_tc_enabled():
    """First block's backward with its heavy part on the tensor core (istgcn_tcn2_bwd_up on the
    aggregated input, csrc/gcn_small.cu); ISTGCN_SMALL_BWD_TC=0 keeps the one-kernel CUDA-core form."""
    return os.environ.get('ISTGCN_SMALL_BWD_TC', '1') != '0'


def _gcn_forward(x, Wc, W2, biasterm, vals, pat, z, s_sum, s_sq, frames, V, K, Cin, Cout, math, keep=None):
    """The fused graph convolution: tcgen05 engine when enabled, mma.sync engine otherwise.
    ``keep``: dict that receives what the first block's tensor-core backward needs (X', per-joint sums)."""
    if _gcn_small_ok(Cin, Cout):       # first block: 3 input channels, CUDA cores, full fp32
        xagg = zsum = None
        if keep is not None:
            xagg = torch.empty(frames * V, 16, device=x.device, dtype=torch.float32)
            zsum = torch.zeros(V, Cout, device=x.device, dtype=torch.float32)
            keep['xagg'], keep['zsum'] = xagg, zsum
        call('gcn_small_fwd', x, Wc, biasterm, vals, pat.dst_ptr, pat.dst_src, pat.dst_id, pat.nnz, z,
             s_sum, s_sq, xagg, zsum, frames, V, K, Cin, Cout)
    elif use_tc():
        W2, bias_k, colsum = W2        # graph_conv_operands: weight rows + the bias-term factors
        W2 = W2.contiguous()
        call('gcn_tc', x, None, None, None, None, None, W2, vals, pat.dst_ptr, pat.dst_src,
             pat.dst_id, pat.nnz, None if bias_k is None else bias_k.contiguous(), colsum, None, z, None,
             s_sum, s_sq, frames, V, K, Cin, W2.shape[1], Cout, 0, 0, 1, 0, 0)
    else:
        call('gcn_fwd', x, Wc, biasterm, vals, pat.dst_ptr, pat.dst_src, pat.dst_id, pat.nnz, None, z,
             s_sum, s_sq, frames, V, K, Cin, Cout, 0, 0, 1, 0, math)


class DataBN(Function):
    """data_bn + both permutes: (N, C, T, V, M) -> (N*M, T, V, C)   [st_gcnold.py:74-80]."""

    @staticmethod
    def forward(ctx, x, weight, bias, st, training):
        x = x.contiguous()
        N, C, T, V, M = x.shape
        dev = x.device
        if training:
            stats = torch.zeros(2, V * C, device=dev, dtype=torch.float64)
            call('data_bn_stats', x, stats[0], stats[1], N, C, T, V, M)
            mean, scale, rstd = _bn_forward_coeffs(True, stats[0], stats[1], N * M * T, weight, st,
                                                   V * C, dev)
        else:
            mean, scale, rstd = _bn_forward_coeffs(False, None, None, 0, weight, st, V * C, dev)
        y = torch.empty(N * M, T, V, C, device=dev, dtype=torch.float32)
        call('data_bn_apply', x, mean, scale, bias.contiguous(), y, N, C, T, V, M)
        ctx.training = training
        ctx.dims = (N, C, T, V, M)
        if training:
            ctx.save_for_backward(x, mean, rstd)
        return y

    @staticmethod
    def backward(ctx, gy):
        if not ctx.training:
            raise RuntimeError('istgcn: backward through eval-mode BatchNorm is not supported')
        x, mean, rstd = ctx.saved_tensors
        N, C, T, V, M = ctx.dims
        d = torch.zeros(2, V * C, device=x.device, dtype=torch.float64)
        call('data_bn_bwd', x, gy.contiguous(), mean, rstd, d[0], d[1], N, C, T, V, M)
        d = d.float()
        return None, d[0], d[1], None, None


class BlockCfg(object):
    """Static (non-tensor) description of one block, built once per module."""

    def __init__(self, pattern, ident, stride, res_mode, drop_p, bn1, bn2, bnr, bp):
        self.pattern, self.ident = pattern, ident
        self.stride, self.res_mode, self.drop_p, self.bp = stride, res_mode, float(drop_p), bp
        self.bn1, self.bn2, self.bnr = bn1, bn2, bnr
        self.ones = None
        self.training = True
        self.seed = 0


def _gcn_tc2_ok(cin, cout):
    """Channel counts the TMA-fed graph-convolution engine takes (csrc/gcn_tc2.cu:
    gcn_tc2_eligible); everything else stays on the first-generation kernel."""
    return cin % 32 == 0 and cout % 32 == 0 and 32 <= cout <= 256 and \
        os.environ.get('ISTGCN_GCN_TC_V1') is None


def _gcn_pair_ok(cin, cout):
    """Shapes the one-pass weight + adjacency gradient takes (csrc/gcn_pair_tc.cu)."""
    return cin % 32 == 0 and cout % 32 == 0 and cin <= 512 and os.environ.get('ISTGCN_GCN_PAIR_OFF') is None


def gcn_pair_grads(dz, x, vals, Wc, pat, dWc, dvals, frames, V, K, Cin, Cout):
    items, ctas, joints = pat.pair_items(Cin, Cout)
    ws = torch.zeros(pat.npairs, Cin, Cout, device=dz.device, dtype=torch.float32)
    call('gcn_pair_grads', dz, x, vals, Wc, items, items.shape[0], ctas, ctas.shape[0], joints, pat.pair_of, pat.npairs,
         pat.entry_pair, pat.k_ptr, pat.nnz, ws, dWc, dvals, frames, V, K, Cin, Cout)


# The weight / adjacency gradient of a block is needed only by the parameter-regrouping backward
# (modules.BlockPrep.backward) at the very end of the backward pass, and its kernel is bound by the
# tensor pipe and L2 while the kernels that follow it on the main stream (the next block's
# element-wise and temporal backward) are bound by HBM: it runs on its own stream, and the consumer
# waits for the event registered under the gradient's address.
_pair_streams = {}
_pair_events = {}


def _pair_async_enabled():
    return os.environ.get('ISTGCN_PAIR_ASYNC', '1') != '0'


def _pair_stream(device):
    key = (device.type, device.index)
    if key not in _pair_streams:
        _pair_streams[key] = torch.cuda.Stream(device=device)
    return _pair_streams[key]


def run_on_grad_stream(key, tensors, fn):
    """Run ``fn()`` (weight-gradient kernels nobody on the main stream waits for) on the gradient
    stream, ordered after everything already on the current stream.  ``tensors`` = everything fn
    touches that was allocated on another stream; ``key`` = the gradient tensor whose consumer calls
    wait_pair_grads(key)."""
    dev = key.device
    main = torch.cuda.current_stream(dev)
    ps = _pair_stream(dev)
    ev0 = torch.cuda.Event()
    ev0.record(main)
    ps.wait_event(ev0)
    with torch.cuda.stream(ps):
        fn()
        ev1 = torch.cuda.Event()
        ev1.record(ps)
    for t in tensors:
        if t is not None:
            t.record_stream(ps)
    _pair_events.setdefault(key.data_ptr(), []).append(ev1)


def gcn_pair_grads_async(dz, x, vals, Wc, pat, dWc, dvals, owner, frames, V, K, Cin, Cout):
    """gcn_pair_grads on the gradient stream.  `owner` is the allocation dWc / dvals are views of."""
    run_on_grad_stream(dWc, (dz, x, vals, Wc, owner),
                       lambda: gcn_pair_grads(dz, x, vals, Wc, pat, dWc, dvals, frames, V, K, Cin, Cout))


def wait_pair_grads(dWc):
    """The current stream waits for every gradient-stream kernel registered under ``dWc`` (no-op if
    there is none)."""
    if dWc is None:
        return
    for ev in _pair_events.pop(dWc.data_ptr(), ()):
        torch.cuda.current_stream(dWc.device).wait_event(ev)


def _gcn_small_ok(cin, cout):
    """The first block's narrow-input graph convolution (csrc/gcn_small.cu)."""
    # fast mode only: the fp32-grade parity mode keeps one arithmetic (the 3xTF32 engine) for every
    # block, which is what its golden gradient bounds were calibrated on
    return cin <= 4 and cout == 64 and use_tc() and os.environ.get('ISTGCN_GCN_SMALL_OFF') is None


def _tcn2_ok(C, bp):
    """Shapes the streaming TCN kernels of the fast mode take (csrc/tcn2.cu); the fp32-grade
    '3xtf32' mode keeps the error-compensated kernels of csrc/tcn.cu."""
    return use_tc() and C in (64, 128, 256) and bp in (8, 16) and \
        os.environ.get('ISTGCN_TCN_V1') is None


def _tconv_fused_ok(C, V):
    """Shapes the tcgen05 temporal-convolution kernel takes (csrc/tconv_tc.cu)."""
    F = min(8, 128 // V)
    return C % 32 == 0 and F * V > 96


class STBlock(Function):
    """One IST-GCN block: graph conv -> BN -> ReLU -> 1x1 -> {3,9,15}x1 -> 1x1 -> BN -> dropout
    -> + residual -> ReLU  (net/st_gcn_mstcn_1x1.py:250-266 with tgcn.py:76-89 or
    inceptionv2_gcn.py:64-89 as the graph conv)."""

    @staticmethod
    def forward(ctx, x, vals, Wc, biasterm, W2, bn1_w, bn1_b, Wd, bd, Weff, beff, Wu, bu, bn2_w,
                bn2_b, Wr, biasterm_r, bnr_w, bnr_b, cfg):
        x = x.contiguous()
        NM, T, V, Cin = x.shape
        Cout = Wc.shape[1]
        pat, s, bp = cfg.pattern, cfg.stride, cfg.bp
        K = pat.K
        Tout = (T - 1) // s + 1
        R_in, R_out = NM * T * V, NM * Tout * V
        dev = x.device
        training = cfg.training
        math = math_flag()
        drop_p = cfg.drop_p if training else 0.0
        vals, Wc, biasterm = vals.contiguous(), Wc.contiguous(), biasterm.contiguous()
        Wd, bd, Weff, beff = Wd.contiguous(), bd.contiguous(), Weff.contiguous(), beff.contiguous()
        Wu, bu = Wu.contiguous(), bu.contiguous()
        stats = torch.zeros(6, Cout, device=dev, dtype=torch.float64) if training else [None] * 6

        z = torch.empty(NM, T, V, Cout, device=dev, dtype=torch.float32)
        tcn2 = _tcn2_ok(Cout, bp)
        # first block, training, fast mode: the backward runs on the tensor core and wants X' and the
        # per-joint sums of z from this pass
        keep = {} if (training and tcn2 and _gcn_small_ok(Cin, Cout) and cfg.res_mode == 0 and Cin <= 3
                      and pat.K <= 4 and _small_bwd_tc_enabled()) else None
        _gcn_forward(x, Wc, W2, biasterm, vals, pat, z, stats[0], stats[1], NM * T, V, K,
                     Cin, Cout, math, keep)
        fold = tcn2 and training and _bn_fold_enabled()      # BN bookkeeping inside the consumers
        if fold:
            scale1, mean1, rstd1, scale2, mean2, rstd2 = _coeffs(6, Cout, dev)
        else:
            mean1, scale1, rstd1 = _bn_forward_coeffs(training, stats[0], stats[1], R_in, bn1_w, cfg.bn1,
                                                      Cout, dev)
        bn1_b, bn2_b = bn1_b.contiguous(), bn2_b.contiguous()
        h1 = torch.empty(NM, T, V, bp, device=dev, dtype=torch.float32)
        h2 = torch.empty(NM, Tout, V, bp, device=dev, dtype=torch.float32)
        u = torch.empty(NM, Tout, V, Cout, device=dev, dtype=torch.float32)
        if tcn2:
            if fold:
                st1 = cfg.bn1
                call('tcn2_down_bn', z, stats[0], stats[1], f64(R_in), bn1_w, bn1_b, st1.running_mean,
                     st1.running_var, float(st1.momentum), float(st1.eps), mean1, scale1, rstd1, Wd, bd, h1,
                     i64(R_in), Cout, bp)
            else:
                call('tcn2_down', z, mean1, scale1, bn1_b, Wd, bd, h1, i64(R_in), Cout, bp)
            call('tcn2_conv', h1, Weff, beff, h2, NM, T, V, bp, s)
            call('tcn2_up', h2, Wu, bu, u, stats[2], stats[3], i64(R_out), Cout, bp)
        else:
            call('tcn_fwd', z, mean1, scale1, bn1_b, Wd, bd, Weff, beff, Wu, bu, h1, h2, u, stats[2],
                 stats[3], NM, T, V, Cout, bp, s, math)
        if not fold:
            mean2, scale2, rstd2 = _bn_forward_coeffs(training, stats[2], stats[3], R_out, bn2_w, cfg.bn2,
                                                      Cout, dev)
        rres = scale_r = mean_r = rstd_r = None
        res = None
        if cfg.res_mode == 1:
            res = x
        elif cfg.res_mode == 2:
            idn = cfg.ident
            if cfg.ones is None or cfg.ones.device != dev:
                cfg.ones = torch.ones(V, device=dev, dtype=torch.float32)
            Wr, biasterm_r = Wr.contiguous(), biasterm_r.contiguous()
            rres = torch.empty(NM, Tout, V, Cout, device=dev, dtype=torch.float32)
            if use_tc() and _tconv_fused_ok(Cin, V) and Cout % 32 == 0:
                # strided 1x1 conv = the one-tap case of the TMA-fed temporal convolution
                call('tconv_tc', x, Wr.t().contiguous(), biasterm_r[0].contiguous(), rres, stats[4],
                     stats[5], NM, T, Tout, V, Cin, Cout, 1, s, 1)
            elif use_tc():   # K=1 / identity-adjacency case of the graph-conv engine
                W2r = Wr.t().contiguous()                      # (Cout, Cin): rows n, cols ci
                call('gcn_tc', x, None, None, None, None, None, W2r, cfg.ones, idn.dst_ptr,
                     idn.dst_src, idn.dst_id, V, biasterm_r[0].contiguous(), cfg.ones, None, rres, None,
                     stats[4], stats[5],
                     NM * Tout, V, 1, Cin, Cin, Cout, T, Tout, s, 0, 1)
            else:
                call('gcn_fwd', x, Wr, biasterm_r, cfg.ones, idn.dst_ptr, idn.dst_src, idn.dst_id, V,
                     None, rres, stats[4], stats[5], NM * Tout, V, 1, Cin, Cout, T, Tout, s, 0, math)
            if fold:
                scale_r, mean_r, rstd_r = _coeffs(3, Cout, dev)
            else:
                mean_r, scale_r, rstd_r = _bn_forward_coeffs(training, stats[4], stats[5], R_out, bnr_w,
                                                             cfg.bnr, Cout, dev)
            bnr_b = bnr_b.contiguous()
            res = rres
        out = torch.empty(NM, Tout, V, Cout, device=dev, dtype=torch.float32)
        if fold:
            st2, str_ = cfg.bn2, cfg.bnr
            has_r = cfg.res_mode == 2
            call('block_tail_fwd_bn', u, stats[2], stats[3], f64(R_out), bn2_w, bn2_b, st2.running_mean,
                 st2.running_var, float(st2.momentum), float(st2.eps), mean2, scale2, rstd2, res,
                 stats[4] if has_r else None, stats[5] if has_r else None, bnr_w if has_r else None,
                 bnr_b if has_r else None, str_.running_mean if has_r else None,
                 str_.running_var if has_r else None, float(str_.momentum) if has_r else 0.0,
                 float(str_.eps) if has_r else 0.0, mean_r, scale_r, rstd_r, out, i64(R_out), Cout,
                 float(drop_p), u64(cfg.seed), step_counter(dev))
        else:
            call('block_tail_fwd', u, mean2, scale2, bn2_b, res, mean_r, scale_r,
                 bnr_b if cfg.res_mode == 2 else None, out, i64(R_out), Cout, float(drop_p),
                 u64(cfg.seed), step_counter(dev))

        ctx.cfg, ctx.training, ctx.math, ctx.drop_p, ctx.seed = cfg, training, math, drop_p, cfg.seed
        ctx.dims = (NM, T, Tout, V, Cin, Cout)
        ctx.tcn2 = tcn2
        ctx.small_keep = keep
        if training:
            ctx.save_for_backward(x, vals, Wc, z, h1, h2, u, out, rres, scale1, bn1_b, mean1, rstd1,
                                  mean2, rstd2, mean_r, rstd_r, Wd, Weff, Wu, Wr, bn1_w, bn2_w, bnr_w)
        return out

    @staticmethod
    def backward(ctx, gout):
        if not ctx.training:
            raise RuntimeError('istgcn: backward through eval-mode BatchNorm is not supported')
        (x, vals, Wc, z, h1, h2, u, out, rres, scale1, beta1, mean1, rstd1, mean2, rstd2, mean_r,
         rstd_r, Wd, Weff, Wu, Wr, bn1_w, bn2_w, bnr_w) = ctx.saved_tensors
        cfg, math, drop_p, seed = ctx.cfg, ctx.math, ctx.drop_p, ctx.seed
        NM, T, Tout, V, Cin, Cout = ctx.dims
        pat, s, bp = cfg.pattern, cfg.stride, cfg.bp
        K = pat.K
        R_in, R_out = NM * T * V, NM * Tout * V
        dev = x.device
        gout = gout.contiguous()
        sums = torch.zeros(6, Cout, device=dev, dtype=torch.float64)
        go = torch.empty_like(gout)
        call('block_tail_bwd', gout, out, u, mean2, rstd2, rres, mean_r, rstd_r, go, sums[0], sums[1],
             sums[2] if rres is not None else None, sums[3] if rres is not None else None,
             i64(R_out), Cout, float(drop_p), u64(seed), step_counter(dev))
        p2, m12, c2, dg2, db2 = _coeffs(5, Cout, dev)
        fold = ctx.tcn2 and _bn_fold_enabled()          # bn_bwd_coeffs inside the consumer kernels
        if not fold:
            call('bn_bwd_coeffs', sums[0], sums[1], f64(R_out), bn2_w, rstd2, p2, m12, c2, dg2, db2, Cout)
        # every caller-zeroed fp32 accumulator of this block out of ONE zero-filled buffer (one fill
        # launch instead of twelve); each view starts on a 16-byte boundary
        need_r = cfg.res_mode == 2
        shapes = [tuple(Wd.shape), (bp,), tuple(Weff.shape), (bp,), tuple(Wu.shape), (Cout,),
                  tuple(vals.shape), tuple(Wc.shape), (V, Cout)]
        if need_r:
            shapes += [tuple(Wr.shape), (V, Cout)]
        sizes = [(int(torch.Size(sh).numel()) + 3) // 4 * 4 for sh in shapes]
        flat = torch.zeros(sum(sizes), device=dev, dtype=torch.float32)
        views, off = [], 0
        for sh, n in zip(shapes, sizes):
            views.append(flat[off:off + torch.Size(sh).numel()].view(sh))
            off += n
        dWd, dbd, dWeff, dbeff, dWu, dbu, dvals, dWc, dbt = views[:9]
        dWr_z, dbtr_z = (views[9], views[10]) if need_r else (None, None)
        g1 = torch.empty(NM, T, V, Cout, device=dev, dtype=torch.float32)
        dh2 = torch.empty(R_out, bp, device=dev, dtype=torch.float32)
        dh1 = torch.empty(R_in, bp, device=dev, dtype=torch.float32)
        if ctx.tcn2:
            if fold:
                call('tcn2_bwd_up_bn', go, u, sums[0], sums[1], f64(R_out), bn2_w, rstd2, p2, m12, c2, dg2, db2,
                     mean2, h2, Wu, dh2, dWu, dbu, dbeff, i64(R_out), Cout, bp, float(drop_p), u64(seed),
                     step_counter(dev))
            else:
                call('tcn2_bwd_up', go, u, p2, m12, c2, mean2, h2, Wu, dh2, dWu, dbu, dbeff, i64(R_out), Cout,
                     bp, float(drop_p), u64(seed), step_counter(dev))
            if getattr(cfg, 'pair_async', False) and _pair_async_enabled() and \
                    os.environ.get('ISTGCN_DW_ASYNC', '1') != '0':
                # the tap-weight gradient is needed by BlockPrep.backward only: gradient stream
                call('tcn2_bwd_conv', dh2, h1, Weff, dh1, None, dbd, NM, T, V, bp, s)
                run_on_grad_stream(dWc, (dh2, h1, Weff, flat),
                                   lambda: call('tcn2_bwd_conv', dh2, h1, Weff, None, dWeff, None, NM, T, V, bp, s))
            else:
                call('tcn2_bwd_conv', dh2, h1, Weff, dh1, dWeff, dbd, NM, T, V, bp, s)
            call('tcn2_bwd_down', dh1, z, mean1, scale1, beta1, rstd1, Wd, g1, dWd, sums[4], sums[5],
                 i64(R_in), Cout, bp)
        else:
            call('tcn_bwd', go, u, p2, m12, c2, mean2, z, scale1, beta1, mean1, rstd1, h1, h2, Wd, Weff, Wu,
                 dh2, dh1, g1, sums[4], sums[5], dWd, dbd, dWeff, dbeff, dWu, dbu, NM, T, V, Cout, bp, s,
                 float(drop_p), u64(seed), step_counter(dev), math)
        p1, m11, c1, dg1, db1 = _coeffs(5, Cout, dev)
        # BN1's coefficients are folded into bn_back_colsum where that kernel is the consumer
        fold1 = fold and use_tc() and _gcn_tc2_ok(Cout, Cin) and not (_gcn_small_ok(Cin, Cout) and cfg.res_mode == 0)
        if not fold1:
            call('bn_bwd_coeffs', sums[4], sums[5], f64(R_in), bn1_w, rstd1, p1, m11, c1, dg1, db1, Cout)
        # identity residual: the block-input gradient starts as `go` and the graph-conv input
        # gradient is accumulated onto it in place (TMA reduce-add on the tcgen05 engine); every
        # other reader of `go` has already run on this stream
        gin = go if cfg.res_mode == 1 else torch.empty_like(x)
        add_in = gin if cfg.res_mode == 1 else None
        # strided-conv residual: its input gradient (one tap of the TMA-fed temporal-convolution
        # kernel, transposed) is written FIRST -- every `s`-th frame of a zeroed gin -- and the
        # graph-conv input gradient is then reduce-added on top
        fused_res = cfg.res_mode == 2 and use_tc() and _tconv_fused_ok(Cout, V) and Cin % 32 == 0 \
            and s <= 2 and _gcn_tc2_ok(Cout, Cin)
        dyr = None
        if fused_res:
            pr, m1r, cr, dgr, dbr = _coeffs(5, Cout, dev)
            call('bn_bwd_coeffs', sums[2], sums[3], f64(R_out), bnr_w, rstd_r, pr, m1r, cr, dgr, dbr,
                 Cout)
            dyr = torch.empty_like(rres)
            call('bn_back_apply', go, rres, pr, m1r, cr, mean_r, dyr, i64(R_out), Cout, 0.0, u64(0), None)
            gin = torch.zeros_like(x)
            call('tconv_tc', dyr, Wr.contiguous(), None, gin, None, None, NM, T, Tout, V, Cout, Cin, 1, s, -1)
            add_in = gin
        small = _gcn_small_ok(Cin, Cout) and cfg.res_mode == 0
        dbt_done = pair = False
        if small and ctx.small_keep is not None:
            # first block, tensor-core form (csrc/gcn_small.cu, second form): the up-projection backward
            # of the temporal chain on (g1, z, X', Wc16) gives G = dz Wc^T and the weight gradient in one
            # pass over (g1, z); per-joint sums of g1; dx / dvals / dbt from the 16-wide G
            keep = ctx.small_keep
            Wc16 = torch.zeros(4, 4, Cout, device=dev, dtype=torch.float32)
            Wc16[:K, :Cin] = Wc.view(K, Cin, Cout)
            G = torch.empty(R_in, 16, device=dev, dtype=torch.float32)
            scratch = torch.zeros(16 * Cout + Cout + 16 + V * Cout, device=dev, dtype=torch.float32)
            dW16 = scratch[:16 * Cout].view(16, Cout)
            sg1 = scratch[16 * Cout + Cout + 16:].view(V, Cout)
            call('tcn2_bwd_up', g1, z, p1, m11, c1, mean1, keep['xagg'], Wc16.view(16, Cout), G, dW16,
                 scratch[16 * Cout:16 * Cout + Cout], scratch[16 * Cout + Cout:16 * Cout + Cout + 16],
                 i64(R_in), Cout, 16, 0.0, u64(0), None)
            call('joint_colsum', g1, sg1, NM * T, V, Cout)
            call('gcn_small_bwd_post', G, x, vals, pat.dst_ptr, pat.dst_src, pat.dst_id, pat.t_ptr, pat.t_src,
                 pat.t_id, pat.nnz, gin, dvals, sg1, keep['zsum'], p1, m11, c1, mean1, dbt, NM * T, V, K, Cin,
                 Cout)
            dWc += dW16.view(4, 4, Cout)[:K, :Cin].reshape(K * Cin, Cout)
        elif small:
            # first block: dz, dx, dvals, dWc and dbt in one CUDA-core kernel
            call('gcn_small_bwd', g1, z, p1, m11, c1, mean1, x, Wc, vals, pat.dst_ptr, pat.dst_src,
                 pat.dst_id, pat.t_ptr, pat.t_src, pat.t_id, pat.nnz, gin, dvals, dWc, dbt, NM * T, V,
                 K, Cin, Cout)
        elif use_tc():
            # input gradient on the tcgen05 engine: the forward kernel run on dz with the
            # transposed adjacency lists and Wc as the weight; adjacency gradient separately
            dz = torch.empty_like(z)
            if _gcn_tc2_ok(Cout, Cin):
                # second-generation engine (csrc/gcn_tc2.cu): its input arrives by TMA, so dz
                # (BatchNorm backward of g1) is materialised by the element-wise kernel first
                # (the same pass sums dz over frames: the bias-term gradient dbt)
                if fold1:
                    call('bn_back_colsum_bn', g1, z, sums[4], sums[5], f64(R_in), bn1_w, rstd1, p1, m11, c1,
                         dg1, db1, mean1, dz, dbt, NM * T, V, Cout)
                else:
                    call('bn_back_colsum', g1, z, p1, m11, c1, mean1, dz, dbt, NM * T, V, Cout)
                dbt_done = True
                call('gcn_tc', dz, None, None, None, None, None, Wc, vals, pat.t_ptr, pat.t_src,
                     pat.t_id, pat.nnz, None, None, add_in, gin, None, None, None, NM * T, V, K, Cout,
                     Cout, Cin, 0, 0, 1, 0, 0)
            else:
                call('gcn_tc', g1, z, p1, m11, c1, mean1, Wc, vals, pat.t_ptr, pat.t_src, pat.t_id,
                     pat.nnz, None, None, add_in, gin, dz, None, None, NM * T, V, K, Cout, Cout, Cin,
                     0, 0, 1, 0, 0)
            pair = dbt_done and _gcn_pair_ok(Cin, Cout)
            if pair and getattr(cfg, 'pair_async', False) and _pair_async_enabled():
                gcn_pair_grads_async(dz, x, vals, Wc, pat, dWc, dvals, flat, NM * T, V, K, Cin, Cout)
            elif pair:    # weight AND adjacency gradient from one pass over (dz, x)
                gcn_pair_grads(dz, x, vals, Wc, pat, dWc, dvals, NM * T, V, K, Cin, Cout)
            else:
                call('gcn_tc_dvals', dz, x, Wc, pat.dst_ptr, pat.dst_src, pat.dst_id, pat.nnz, dvals,
                     NM * T, V, K, Cin, Cout)
        else:
            call('gcn_bwd_x', g1, z, p1, m11, c1, mean1, x, Wc, vals, pat.src_ptr, pat.src_kw,
                 pat.src_id, pat.nnz, add_in, gin, dvals, NM * T, V, K, Cin, Cout, 0, 0, 1, 0, math)
        if small:
            pass
        elif use_tc() and pair:
            pass
        elif use_tc():
            call('gcn_tc_dw', dz, x, vals, pat.dst_ptr, pat.dst_src, pat.dst_id, pat.nnz, dWc,
                 None if dbt_done else dbt, NM * T, V, K, Cin, Cout, 0, 0, 1, 0)
        else:
            call('gcn_bwd_w', g1, z, p1, m11, c1, mean1, x, vals, pat.dst_ptr, pat.dst_src,
                 pat.dst_id, pat.nnz, dWc, dbt, NM * T, V, K, Cin, Cout, 0, 0, 1, 0, math)
        dWr = dbtr = None
        if not fused_res:
            dgr = dbr = None
        if fused_res:
            dWr, dbtr = dWr_z, dbtr_z
            if Cout <= 128 or Cout % 128 == 0:
                call('tconv_dw_tc', x, dyr, dWr, dbtr, NM, T, Tout, V, Cin, Cout, 1, s)
            else:
                idn = cfg.ident
                call('gcn_tc_dw', dyr, x, cfg.ones, idn.dst_ptr, idn.dst_src, idn.dst_id, V, dWr,
                     dbtr, NM * Tout, V, 1, Cin, Cout, T, Tout, s, 0)
        elif cfg.res_mode == 2:
            idn = cfg.ident
            pr, m1r, cr, dgr, dbr = _coeffs(5, Cout, dev)
            call('bn_bwd_coeffs', sums[2], sums[3], f64(R_out), bnr_w, rstd_r, pr, m1r, cr, dgr, dbr,
                 Cout)
            dWr, dbtr = dWr_z, dbtr_z
            if use_tc():
                dyr = torch.empty_like(rres)
                call('gcn_tc', go, rres, pr, m1r, cr, mean_r, Wr, cfg.ones, idn.t_ptr, idn.t_src,
                     idn.t_id, V, None, None, gin, gin, dyr, None, None, NM * Tout, V, 1, Cout, Cout, Cin,
                     T, Tout, s, 0, 2)
                if Cin % 32 == 0 and (Cout <= 128 or Cout % 128 == 0):
                    call('tconv_dw_tc', x, dyr, dWr, dbtr, NM, T, Tout, V, Cin, Cout, 1, s)
                else:
                    call('gcn_tc_dw', dyr, x, cfg.ones, idn.dst_ptr, idn.dst_src, idn.dst_id, V, dWr,
                         dbtr, NM * Tout, V, 1, Cin, Cout, T, Tout, s, 0)
            else:
                call('gcn_bwd_x', go, rres, pr, m1r, cr, mean_r, x, Wr, cfg.ones, idn.src_ptr,
                     idn.src_kw, idn.src_id, V, gin, gin, None, NM * Tout, V, 1, Cin, Cout, T, Tout, s,
                     0, math)
                call('gcn_bwd_w', go, rres, pr, m1r, cr, mean_r, x, cfg.ones, idn.dst_ptr, idn.dst_src,
                     idn.dst_id, V, dWr, dbtr, NM * Tout, V, 1, Cin, Cout, T, Tout, s, 0, math)
        return (gin, dvals, dWc, dbt, None, dg1, db1, dWd, dbd, dWeff, dbeff, dWu, dbu, dg2, db2, dWr,
                dbtr, dgr, dbr, None)


class STBlockWide(Function):
    """One baseline ST-GCN block with a FULL-WIDTH temporal convolution: graph conv -> BN -> ReLU
    -> Conv2d(C, C, (kt,1), (stride,1), (pad,0)) -> BN -> dropout -> + residual -> ReLU
    (net/st_gcnold.py:148-203; net/st_gcn_msgcn.py with the Inception graph conv;
    net/st_gcn_mstcn.py with the three 3/9/15-tap branches merged into one 15-tap kernel).

    The temporal convolution is the sum over taps of the strided, shifted 1x1 engine (K = 1,
    identity adjacency, t_offset = tap - pad); its partial sums accumulate in place.
    Wtt[kt, C(in), C(out)] is the conv weight with the tap first, bt[C] its bias."""

    @staticmethod
    def forward(ctx, x, vals, Wc, biasterm, W2, bn1_w, bn1_b, Wtt, bt, bn2_w, bn2_b, Wr, biasterm_r,
                bnr_w, bnr_b, cfg):
        x = x.contiguous()
        NM, T, V, Cin = x.shape
        C = Wc.shape[1]
        pat, s, idn = cfg.pattern, cfg.stride, cfg.ident
        K, kt = pat.K, Wtt.shape[0]
        pad = (kt - 1) // 2
        Tout = (T - 1) // s + 1
        R_in, R_out = NM * T * V, NM * Tout * V
        dev = x.device
        training = cfg.training
        math = math_flag()
        drop_p = cfg.drop_p if training else 0.0
        vals, Wc, biasterm = vals.contiguous(), Wc.contiguous(), biasterm.contiguous()
        Wtt, bt = Wtt.contiguous(), bt.contiguous()
        if cfg.ones is None or cfg.ones.device != dev:
            cfg.ones = torch.ones(V, device=dev, dtype=torch.float32)
        stats = torch.zeros(6, C, device=dev, dtype=torch.float64) if training else [None] * 6

        z = torch.empty(NM, T, V, C, device=dev, dtype=torch.float32)
        _gcn_forward(x, Wc, W2, biasterm, vals, pat, z, stats[0], stats[1], NM * T, V, K, Cin, C, math)
        mean1, scale1, rstd1 = _bn_forward_coeffs(training, stats[0], stats[1], R_in, bn1_w, cfg.bn1,
                                                  C, dev)
        bn1_b, bn2_b = bn1_b.contiguous(), bn2_b.contiguous()
        a = torch.empty_like(z)
        call('bn_relu_apply', z, mean1, scale1, bn1_b, a, i64(R_in), C)
        u = torch.empty(NM, Tout, V, C, device=dev, dtype=torch.float32)
        fused = use_tc() and _tconv_fused_ok(C, V)
        if use_tc():
            Wrows = Wtt.detach().transpose(1, 2).contiguous()         # [kt][C(out)][C(in)]
            bt_k = bt.detach().view(1, C)
        else:
            bt_vc = bt.detach().unsqueeze(0).expand(V, C).contiguous()
        if fused:       # one implicit GEMM: taps = shifted TMA boxes of the activation
            call('tconv_tc', a, Wrows, bt.detach(), u, stats[2], stats[3], NM, T, Tout, V, C, C, kt,
                 s, 1)
        for tap in range(0 if fused else kt):
            first, last = tap == 0, tap == kt - 1
            ssum, ssq = (stats[2], stats[3]) if last else (None, None)
            if use_tc():
                call('gcn_tc', a, None, None, None, None, None, Wrows[tap], cfg.ones, idn.dst_ptr,
                     idn.dst_src, idn.dst_id, V, bt_k if first else None, cfg.ones if first else None,
                     None if first else u, u, None, ssum, ssq, NM * Tout, V, 1, C, C, C, T, Tout, s,
                     tap - pad, 1)
            else:
                call('gcn_fwd', a, Wtt[tap], bt_vc if first else None, cfg.ones, idn.dst_ptr,
                     idn.dst_src, idn.dst_id, V, None if first else u, u, ssum, ssq, NM * Tout, V, 1,
                     C, C, T, Tout, s, tap - pad, math)
        mean2, scale2, rstd2 = _bn_forward_coeffs(training, stats[2], stats[3], R_out, bn2_w, cfg.bn2,
                                                  C, dev)
        rres = scale_r = mean_r = rstd_r = None
        res = None
        if cfg.res_mode == 1:
            res = x
        elif cfg.res_mode == 2:
            Wr, biasterm_r = Wr.contiguous(), biasterm_r.contiguous()
            rres = torch.empty(NM, Tout, V, C, device=dev, dtype=torch.float32)
            if use_tc() and _tconv_fused_ok(Cin, V) and C % 32 == 0:
                call('tconv_tc', x, Wr.t().contiguous(), biasterm_r[0].contiguous(), rres, stats[4],
                     stats[5], NM, T, Tout, V, Cin, C, 1, s, 1)
            elif use_tc():
                W2r = Wr.t().contiguous()
                call('gcn_tc', x, None, None, None, None, None, W2r, cfg.ones, idn.dst_ptr,
                     idn.dst_src, idn.dst_id, V, biasterm_r[0].contiguous(), cfg.ones, None, rres, None,
                     stats[4], stats[5], NM * Tout, V, 1, Cin, Cin, C, T, Tout, s, 0, 1)
            else:
                call('gcn_fwd', x, Wr, biasterm_r, cfg.ones, idn.dst_ptr, idn.dst_src, idn.dst_id, V,
                     None, rres, stats[4], stats[5], NM * Tout, V, 1, Cin, C, T, Tout, s, 0, math)
            mean_r, scale_r, rstd_r = _bn_forward_coeffs(training, stats[4], stats[5], R_out, bnr_w,
                                                         cfg.bnr, C, dev)
            bnr_b = bnr_b.contiguous()
            res = rres
        out = torch.empty(NM, Tout, V, C, device=dev, dtype=torch.float32)
        call('block_tail_fwd', u, mean2, scale2, bn2_b, res, mean_r, scale_r,
             bnr_b if cfg.res_mode == 2 else None, out, i64(R_out), C, float(drop_p),
             u64(cfg.seed), step_counter(dev))

        ctx.cfg, ctx.training, ctx.math, ctx.drop_p, ctx.seed = cfg, training, math, drop_p, cfg.seed
        ctx.dims = (NM, T, Tout, V, Cin, C)
        if training:
            ctx.save_for_backward(x, vals, Wc, z, a, u, out, rres, mean1, rstd1, mean2, rstd2, mean_r,
                                  rstd_r, Wtt, Wr, bn1_w, bn2_w, bnr_w)
        return out

    @staticmethod
    def backward(ctx, gout):
        if not ctx.training:
            raise RuntimeError('istgcn: backward through eval-mode BatchNorm is not supported')
        (x, vals, Wc, z, a, u, out, rres, mean1, rstd1, mean2, rstd2, mean_r, rstd_r, Wtt, Wr, bn1_w,
         bn2_w, bnr_w) = ctx.saved_tensors
        cfg, math, drop_p, seed = ctx.cfg, ctx.math, ctx.drop_p, ctx.seed
        NM, T, Tout, V, Cin, C = ctx.dims
        pat, s, idn = cfg.pattern, cfg.stride, cfg.ident
        K, kt = pat.K, Wtt.shape[0]
        pad = (kt - 1) // 2
        R_in, R_out = NM * T * V, NM * Tout * V
        dev = x.device
        gout = gout.contiguous()
        sums = torch.zeros(6, C, device=dev, dtype=torch.float64)
        go = torch.empty_like(gout)
        call('block_tail_bwd', gout, out, u, mean2, rstd2, rres, mean_r, rstd_r, go, sums[0], sums[1],
             sums[2] if rres is not None else None, sums[3] if rres is not None else None,
             i64(R_out), C, float(drop_p), u64(seed), step_counter(dev))
        p2, m12, c2, dg2, db2 = _coeffs(5, C, dev)
        call('bn_bwd_coeffs', sums[0], sums[1], f64(R_out), bn2_w, rstd2, p2, m12, c2, dg2, db2, C)
        du = torch.empty_like(u)
        call('bn_back_apply', go, u, p2, m12, c2, mean2, du, i64(R_out), C, float(drop_p), u64(seed),
             step_counter(dev))
        # gradient w.r.t. a = relu(BN1(z)): transposed taps, accumulated; weight / bias gradients
        da = torch.zeros(NM, T, V, C, device=dev, dtype=torch.float32)
        dWtt, dbt_vc = torch.zeros_like(Wtt), torch.zeros(V, C, device=dev)
        # transposed convolution = the same implicit GEMM with mirrored taps (stride 2: one launch
        # per input-frame parity inside the entry point)
        fused_dx = use_tc() and _tconv_fused_ok(C, V) and s <= 2
        if fused_dx:
            call('tconv_tc', du, Wtt, None, da, None, None, NM, T, Tout, V, C, C, kt, s, -1)
        fused_dw = use_tc() and C % 32 == 0 and (C <= 128 or C % 128 == 0)
        if fused_dw:    # all taps of the weight gradient in one kernel (accumulators in TMEM)
            call('tconv_dw_tc', a, du, dWtt, dbt_vc, NM, T, Tout, V, C, C, kt, s)
        for tap in range(kt):
            off = tap - pad
            if use_tc():
                if fused_dx:
                    pass
                elif s == 1:      # pure shift: read du at frame t - off, whole tiles leave through TMA
                    call('gcn_tc', du, None, None, None, None, None, Wtt[tap], cfg.ones, idn.t_ptr,
                         idn.t_src, idn.t_id, V, None, None, None if tap == 0 else da, da, None, None,
                         None, NM * T, V, 1, C, C, C, T, T, 1, -off, 1)
                else:
                    call('gcn_tc', du, None, None, None, None, None, Wtt[tap], cfg.ones, idn.t_ptr,
                         idn.t_src, idn.t_id, V, None, None, da, da, None, None, None, NM * Tout, V, 1,
                         C, C, C, T, Tout, s, off, 2)
                if not fused_dw:
                    call('gcn_tc_dw', du, a, cfg.ones, idn.dst_ptr, idn.dst_src, idn.dst_id, V, dWtt[tap],
                         dbt_vc if tap == 0 else None, NM * Tout, V, 1, C, C, T, Tout, s, off)
            else:
                call('gcn_bwd_x', du, None, None, None, None, None, a, Wtt[tap], cfg.ones, idn.src_ptr,
                     idn.src_kw, idn.src_id, V, da, da, None, NM * Tout, V, 1, C, C, T, Tout, s, off,
                     math)
                call('gcn_bwd_w', du, None, None, None, None, None, a, cfg.ones, idn.dst_ptr,
                     idn.dst_src, idn.dst_id, V, dWtt[tap], dbt_vc if tap == 0 else None, NM * Tout, V,
                     1, C, C, T, Tout, s, off, math)
        g1 = torch.empty_like(z)
        call('relu_bn_bwd', da, a, z, mean1, rstd1, g1, sums[4], sums[5], i64(R_in), C)
        p1, m11, c1, dg1, db1 = _coeffs(5, C, dev)
        call('bn_bwd_coeffs', sums[4], sums[5], f64(R_in), bn1_w, rstd1, p1, m11, c1, dg1, db1, C)
        # identity residual: the block-input gradient starts as `go` and the graph-conv input
        # gradient is accumulated onto it in place (TMA reduce-add on the tcgen05 engine); every
        # other reader of `go` has already run on this stream
        gin = go if cfg.res_mode == 1 else torch.empty_like(x)
        dvals = torch.zeros_like(vals)
        add_in = gin if cfg.res_mode == 1 else None
        dWc, dbt = torch.zeros_like(Wc), torch.zeros(V, C, device=dev)
        if use_tc():
            dz = torch.empty_like(z)
            if _gcn_tc2_ok(C, Cin):
                call('bn_back_apply', g1, z, p1, m11, c1, mean1, dz, i64(R_in), C, 0.0, u64(0), None)
                call('gcn_tc', dz, None, None, None, None, None, Wc, vals, pat.t_ptr, pat.t_src,
                     pat.t_id, pat.nnz, None, None, add_in, gin, None, None, None, NM * T, V, K, C, C,
                     Cin, 0, 0, 1, 0, 0)
            else:
                call('gcn_tc', g1, z, p1, m11, c1, mean1, Wc, vals, pat.t_ptr, pat.t_src, pat.t_id,
                     pat.nnz, None, None, add_in, gin, dz, None, None, NM * T, V, K, C, C, Cin, 0, 0, 1,
                     0, 0)
            call('gcn_tc_dvals', dz, x, Wc, pat.dst_ptr, pat.dst_src, pat.dst_id, pat.nnz, dvals,
                 NM * T, V, K, Cin, C)
            call('gcn_tc_dw', dz, x, vals, pat.dst_ptr, pat.dst_src, pat.dst_id, pat.nnz, dWc, dbt,
                 NM * T, V, K, Cin, C, 0, 0, 1, 0)
        else:
            call('gcn_bwd_x', g1, z, p1, m11, c1, mean1, x, Wc, vals, pat.src_ptr, pat.src_kw,
                 pat.src_id, pat.nnz, add_in, gin, dvals, NM * T, V, K, Cin, C, 0, 0, 1, 0, math)
            call('gcn_bwd_w', g1, z, p1, m11, c1, mean1, x, vals, pat.dst_ptr, pat.dst_src,
                 pat.dst_id, pat.nnz, dWc, dbt, NM * T, V, K, Cin, C, 0, 0, 1, 0, math)
        dWr = dbtr = dgr = dbr = None
        if cfg.res_mode == 2:
            pr, m1r, cr, dgr, dbr = _coeffs(5, C, dev)
            call('bn_bwd_coeffs', sums[2], sums[3], f64(R_out), bnr_w, rstd_r, pr, m1r, cr, dgr, dbr, C)
            dWr, dbtr = torch.zeros_like(Wr), torch.zeros(V, C, device=dev)
            if use_tc():
                dyr = torch.empty_like(rres)
                call('gcn_tc', go, rres, pr, m1r, cr, mean_r, Wr, cfg.ones, idn.t_ptr, idn.t_src,
                     idn.t_id, V, None, None, gin, gin, dyr, None, None, NM * Tout, V, 1, C, C, Cin,
                     T, Tout, s, 0, 2)
                if Cin % 32 == 0 and (C <= 128 or C % 128 == 0):
                    call('tconv_dw_tc', x, dyr, dWr, dbtr, NM, T, Tout, V, Cin, C, 1, s)
                else:
                    call('gcn_tc_dw', dyr, x, cfg.ones, idn.dst_ptr, idn.dst_src, idn.dst_id, V, dWr,
                         dbtr, NM * Tout, V, 1, Cin, C, T, Tout, s, 0)
            else:
                call('gcn_bwd_x', go, rres, pr, m1r, cr, mean_r, x, Wr, cfg.ones, idn.src_ptr,
                     idn.src_kw, idn.src_id, V, gin, gin, None, NM * Tout, V, 1, Cin, C, T, Tout, s,
                     0, math)
                call('gcn_bwd_w', go, rres, pr, m1r, cr, mean_r, x, cfg.ones, idn.dst_ptr, idn.dst_src,
                     idn.dst_id, V, dWr, dbtr, NM * Tout, V, 1, Cin, C, T, Tout, s, 0, math)
        return (gin, dvals, dWc, dbt, None, dg1, db1, dWtt, dbt_vc.sum(0), dg2, db2, dWr, dbtr, dgr,
                dbr, None)


class GraphConv(Function):
    """The fused graph convolution on its own (no BatchNorm behind it): tgcn.py:76-89 /
    inceptionv2_gcn.py:64-89.  x (NM, T, V, Cin) channels-last -> (NM, T, V, Cout)."""

    @staticmethod
    def forward(ctx, x, vals, Wc, biasterm, W2, pattern):
        x, vals, Wc, biasterm = x.contiguous(), vals.contiguous(), Wc.contiguous(), biasterm.contiguous()
        NM, T, V, Cin = x.shape
        Cout = Wc.shape[1]
        z = torch.empty(NM, T, V, Cout, device=x.device, dtype=torch.float32)
        math = math_flag()
        _gcn_forward(x, Wc, W2, biasterm, vals, pattern, z, None, None, NM * T, V,
                     pattern.K, Cin, Cout, math)
        ctx.pattern, ctx.math = pattern, math
        ctx.save_for_backward(x, vals, Wc)
        return z

    @staticmethod
    def backward(ctx, gz):
        x, vals, Wc = ctx.saved_tensors
        pat, math = ctx.pattern, ctx.math
        NM, T, V, Cin = x.shape
        Cout = Wc.shape[1]
        gz = gz.contiguous()
        gin, dvals = torch.empty_like(x), torch.zeros_like(vals)
        call('gcn_bwd_x', gz, None, None, None, None, None, x, Wc, vals, pat.src_ptr, pat.src_kw, pat.src_id,
             pat.nnz, None, gin, dvals, NM * T, V, pat.K, Cin, Cout, 0, 0, 1, 0, math)
        dWc, dbt = torch.zeros_like(Wc), torch.zeros(V, Cout, device=x.device)
        call('gcn_bwd_w', gz, None, None, None, None, None, x, vals, pat.dst_ptr, pat.dst_src, pat.dst_id,
             pat.nnz, dWc, dbt, NM * T, V, pat.K, Cin, Cout, 0, 0, 1, 0, math)
        return gin, dvals, dWc, dbt, None, None


class Pool(Function):
    """F.avg_pool2d over (T, V) then mean over M (st_gcnold.py:89-90): (N*M, T, V, C) -> (N, C)."""

    @staticmethod
    def forward(ctx, x, N, M):
        x = x.contiguous()
        NM, T, V, C = x.shape
        pooled = torch.empty(N, C, device=x.device, dtype=torch.float32)
        call('pool_fwd', x, pooled, N, M, T * V, C)
        ctx.dims = (N, M, T, V, C)
        return pooled

    @staticmethod
    def backward(ctx, g):
        N, M, T, V, C = ctx.dims
        gx = torch.empty(N * M, T, V, C, device=g.device, dtype=torch.float32)
        call('pool_bwd', g.contiguous(), gx, N, M, T * V, C)
        return gx, None, None


def dropout_mask(numel, p, seed, device):
    """The keep-mask STBlock's counter-based dropout uses for a (rows, C) tensor of ``numel``
    elements and ``seed`` -- exported for parity tests."""
    mask = torch.empty(numel, device=device, dtype=torch.uint8)
    call('dropout_mask', mask, _lib.i64(numel), float(p), u64(seed), step_counter(torch.device(device)))
    return mask
