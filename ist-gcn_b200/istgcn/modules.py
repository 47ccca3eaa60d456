"""nn.Module shells shared by the drop-in ``net.*`` models.

The classes here own *no* parameters of their own naming: every ``net/<file>.py`` declares its
sub-modules itself, in the reference's registration order, so ``state_dict()`` has the
reference's keys, order and shapes.  What lives here is the part that is
identical for every variant: turning the reference-layout parameters into the operands of the
fused kernels (tiny differentiable torch ops) and driving ``ops.STBlock``.
"""
import itertools

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .sparse import SparsePattern

_TAPS = 15
_seed_counter = itertools.count(1)


def padded_bottleneck(b):
    """Bottleneck width b = int(sqrt(C)) in {8, 11, 16} -> kernel width bp in {8, 16}."""
    if b <= 8:
        return 8
    if b <= 16:
        return 16
    raise NotImplementedError('bottleneck width %d > 16 is not supported by the TCN kernels' % b)


def to_channels_last(x):
    """(N, C, T, V) -> (N, T, V, C) contiguous."""
    return x.permute(0, 2, 3, 1).contiguous()


def to_channels_first(x):
    """(N, T, V, C) -> (N, C, T, V) contiguous."""
    return x.permute(0, 3, 1, 2).contiguous()


def graph_conv_operands(weight, bias, adjs, pattern):
    """Reference-layout conv weight (K*Cout, Cin, 1, 1) + bias and the adjacency stacks
    [A*imp (, A2*imp2, A3*imp3)] -> (vals[nnz], Wc[K*Cin, Cout], biasterm[V, Cout], (W2, bias_k, colsum))
    where W2[K*Cout, CinPad] is the weight in its own row order with the input channels zero-padded to
    a multiple of 32 (the tcgen05 engine's TMA operand; no gradient flows through it, Wc carries
    the weight gradient).

    Three einsums with A, A2, A3 are one with their sum (inceptionv2_gcn.py:69-88); the conv
    bias passes through the aggregation as bias[k, c] * colsum(A_eff[k])[w]."""
    a_eff = adjs[0]
    for extra in adjs[1:]:
        a_eff = a_eff + extra
    K = a_eff.shape[0]
    kc, cin = weight.shape[0], weight.shape[1]
    cout = kc // K
    vals = a_eff.reshape(-1).index_select(0, pattern.flat_idx)
    wc = weight.view(K, cout, cin).permute(0, 2, 1).reshape(K * cin, cout)
    if bias is None:
        biasterm = a_eff.new_zeros(a_eff.shape[1], cout)
    else:
        biasterm = torch.einsum('kc,kw->wc', bias.view(K, cout), a_eff.sum(1))
    w2 = weight.detach().view(kc, cin)
    if cin % 32:
        w2 = F.pad(w2, (0, 32 - cin % 32))
    if bias is None:
        tc_ops = (w2, None, None)
    else:       # the tcgen05 epilogue rebuilds the bias term from its two factors
        tc_ops = (w2, bias.detach().view(K, cout), a_eff.detach().sum(1).contiguous())
    return vals, wc, biasterm, tc_ops


def bottleneck_tcn_operands(conv_start, tcn_1, tcn_2, tcn_3, conv_end, m_imp):
    """conv_1x1_start / tcn_1,2,3 / conv_1x1_end (st_gcn_mstcn_1x1.py:190-224) ->
    Wd[C, bp], bd[bp], Weff[15, bp, bp] (tap, in, out), beff[bp], Wu[bp, C], bu[C].

    The three branches scaled by mstcn_importance and summed (:257-261) are ONE 15-tap
    convolution: the 3- and 9-tap kernels are centred inside the 15-tap window."""
    b, C = conv_start.weight.shape[0], conv_start.weight.shape[1]
    bp = padded_bottleneck(b)
    wd = F.pad(conv_start.weight.view(b, C).t(), (0, bp - b))
    bd = F.pad(conv_start.bias, (0, bp - b))
    weff, beff = 0, 0
    for i, conv in enumerate((tcn_1, tcn_2, tcn_3)):
        k = conv.weight.shape[2]
        off = (_TAPS - k) // 2
        w = conv.weight[:, :, :, 0].permute(2, 1, 0)            # (tap, in, out)
        weff = weff + F.pad(w, (0, 0, 0, 0, off, off)) * m_imp[i]
        beff = beff + conv.bias * m_imp[i]
    weff = F.pad(weff, (0, bp - b, 0, bp - b))
    beff = F.pad(beff, (0, bp - b))
    wu = F.pad(conv_end.weight.view(C, b).t(), (0, 0, 0, bp - b))
    return wd, bd, weff, beff, wu, conv_end.bias


class BlockPrep(torch.autograd.Function):
    """graph_conv_operands + bottleneck_tcn_operands (+ the residual conv regrouping) of one block as
    ONE kernel forward and ONE backward (istgcn_block_prep_fwd / _bwd, csrc/train.cu) instead of
    ~50 tiny ATen launches each way.  Inputs are the reference-layout parameters themselves."""

    @staticmethod
    def forward(ctx, W, bias, imp1, imp2, imp3, Ws, bs, W1, b1, W2, b2, W3, b3, We, m_imp, Wres, bres,
                bufs, pattern, dims):
        from ._lib import call
        K, V, Cin, Cout, b, bp = dims
        dev = W.device
        A1, A2, A3 = bufs
        new = lambda *sh: torch.empty(*sh, device=dev, dtype=torch.float32)          # noqa: E731
        vals, colsum, Wc, biasterm = new(pattern.nnz), new(K, V), new(K * Cin, Cout), new(V, Cout)
        Wd, bd, Weff, beff, Wu = new(Cout, bp), new(bp), new(15, bp, bp), new(bp), new(bp, Cout)
        Wr = btr = None
        if Wres is not None:
            Wr, btr = new(Cin, Cout), new(V, Cout)
        c = lambda t: None if t is None else t.contiguous()                           # noqa: E731
        call('block_prep_fwd', c(W), c(bias), A1, c(imp1), A2, c(imp2), A3, c(imp3), pattern.flat_idx,
             pattern.dst_ptr, pattern.dst_id, pattern.nnz, vals, colsum, Wc, biasterm, c(Ws), c(bs), c(W1),
             c(b1), c(W2), c(b2), c(W3), c(b3), c(We), c(m_imp), Wd, bd, Weff, beff, Wu, c(Wres), c(bres), Wr,
             btr, K, V, Cin, Cout, b, bp)
        ctx.dims, ctx.bufs, ctx.pattern = dims, bufs, pattern
        ctx.has = (bias is not None, imp1 is not None, imp2 is not None, imp3 is not None, Wres is not None)
        ctx.save_for_backward(bias, colsum, W1, b1, W2, b2, W3, b3, m_imp)
        ctx.mark_non_differentiable(colsum)
        if Wres is None:
            Wr, btr = W.new_zeros(0), W.new_zeros(0)
            ctx.mark_non_differentiable(Wr, btr)
        return vals, Wc, biasterm, colsum, Wd, bd, Weff, beff, Wu, Wr, btr

    @staticmethod
    def backward(ctx, dvals, dWc, dbt, _dcolsum, dWd, dbd, dWeff, dbeff, dWu, dWr, dbtr):
        from ._lib import call
        ops.wait_pair_grads(dWc)          # the block's weight / adjacency gradient ran on its own stream
        bias, colsum, W1, b1, W2, b2, W3, b3, m_imp = ctx.saved_tensors
        K, V, Cin, Cout, b, bp = ctx.dims
        A1, A2, A3 = ctx.bufs
        pat = ctx.pattern
        has_bias, h1, h2, h3, has_res = ctx.has
        dev = colsum.device
        z = lambda g, *sh: torch.zeros(*sh, device=dev) if g is None else g.contiguous()     # noqa: E731
        dvals, dWc, dbt = z(dvals, pat.nnz), z(dWc, K * Cin, Cout), z(dbt, V, Cout)
        dWd, dbd, dWeff, dbeff, dWu = z(dWd, Cout, bp), z(dbd, bp), z(dWeff, 15, bp, bp), z(dbeff, bp), z(dWu, bp, Cout)
        new = lambda *sh: torch.empty(*sh, device=dev, dtype=torch.float32)          # noqa: E731
        dW = new(K * Cout, Cin, 1, 1)
        dbias = new(K * Cout) if has_bias else None
        dimp = [new(K, V, V) if h else None for h in (h1, h2, h3)]
        dWs, dbs, dWe, dm = new(b, Cout, 1, 1), new(b), new(Cout, b, 1, 1), new(3)
        dWt = [new(b, b, kt, 1) for kt in (3, 9, 15)]
        dbt3 = [new(b) for _ in range(3)]
        dWres = dbres = None
        if has_res:
            dWr, dbtr = z(dWr, Cin, Cout), z(dbtr, V, Cout)
            dWres, dbres = new(Cout, Cin, 1, 1), new(Cout)
        else:
            dWr = dbtr = None
        call('block_prep_bwd', dvals, dWc, dbt, dWd, dbd, dWeff, dbeff, dWu, dWr, dbtr, bias if has_bias else None,
             colsum, A1, A2, A3, pat.inv_idx, pat.id_kw, W1.contiguous(), b1.contiguous(), W2.contiguous(),
             b2.contiguous(), W3.contiguous(), b3.contiguous(), m_imp.contiguous(), dW, dbias, dimp[0], dimp[1],
             dimp[2], dWs, dbs, dWt[0], dbt3[0], dWt[1], dbt3[1], dWt[2], dbt3[2], dWe, dm, dWres, dbres, K, V,
             Cin, Cout, b, bp)
        return (dW, dbias, dimp[0], dimp[1], dimp[2], dWs, dbs, dWt[0], dbt3[0], dWt[1], dbt3[1], dWt[2],
                dbt3[2], dWe, dm, dWres, dbres, None, None, None)


def _prep_kernel_enabled():
    import os
    return os.environ.get('ISTGCN_PREP_KERNEL', '1') != '0'


class FusedBlockMixin(object):
    """Drives ops.STBlock for a block whose sub-modules follow the reference's names:
    ``gcn`` (ConvTemporalGraphical or Inception2), ``tcn_start``, ``conv_1x1_start``, ``tcn_1``,
    ``tcn_2``, ``tcn_3``, ``conv_1x1_end``, ``tcn_end``, ``residual``."""

    def _init_fused(self, in_channels, out_channels, stride, dropout, residual):
        self._io = (in_channels, out_channels, stride)
        self._drop_p = float(dropout)
        if not residual:
            self._res_mode = 0
        elif in_channels == out_channels and stride == 1:
            self._res_mode = 1
        else:
            self._res_mode = 2
        self._cfg = None

    def _gcn_conv(self):
        gcn = self.gcn
        return gcn.branch.conv if hasattr(gcn, 'branch') else gcn.conv

    def _block_cfg(self, pattern):
        if self._cfg is None or self._cfg.pattern is not pattern:
            bnr = ops.BNState(self.residual[1]) if self._res_mode == 2 else None
            ident = SparsePattern.identity(pattern.V, pattern.flat_idx.device) \
                if self._res_mode == 2 else None
            b = self.conv_1x1_start.weight.shape[0]
            self._cfg = ops.BlockCfg(pattern, ident, self._io[2], self._res_mode, self._drop_p,
                                     ops.BNState(self.tcn_start[0]), ops.BNState(self.tcn_end[0]),
                                     bnr, padded_bottleneck(b))
        else:                                  # buffers may have been re-assigned by .to()/load
            self._cfg.bn1 = ops.BNState(self.tcn_start[0])
            self._cfg.bn2 = ops.BNState(self.tcn_end[0])
            if self._res_mode == 2:
                self._cfg.bnr = ops.BNState(self.residual[1])
        return self._cfg

    def prepare_operands(self, adjs, m_imp, pattern):
        """The kernel operands of this block from the reference-layout parameters (~50 tiny
        differentiable torch ops).  They depend on parameters only, so the model runs them for
        ALL blocks on a side stream ahead of the block kernels (FusedModelMixin._trunk); autograd
        runs their backward on that stream too, off the critical path."""
        conv = self._gcn_conv()
        self._prep_fused = False
        vals, wc, biasterm, w2 = graph_conv_operands(conv.weight, conv.bias, adjs, pattern)
        wd, bd, weff, beff, wu, bu = bottleneck_tcn_operands(
            self.conv_1x1_start, self.tcn_1, self.tcn_2, self.tcn_3, self.conv_1x1_end, m_imp)
        wr = btr = None
        if self._res_mode == 2:
            rconv = self.residual[0]
            cout, cin = rconv.weight.shape[0], rconv.weight.shape[1]
            wr = rconv.weight.view(cout, cin).t().contiguous()
            btr = rconv.bias.unsqueeze(0).expand(pattern.V, cout).contiguous()
        # everything the kernels touch is made contiguous here, not on the main stream
        w2 = tuple(None if t is None else t.contiguous() for t in w2)
        return (vals.contiguous(), wc.contiguous(), biasterm.contiguous(), w2, wd.contiguous(),
                bd.contiguous(), weff.contiguous(), beff.contiguous(), wu.contiguous(), bu.contiguous(),
                wr, btr)

    def prepare_operands_fused(self, bufs, imps, m_imp, pattern):
        """The same operands from the RAW adjacency buffers (A, A2, A3 or A alone) and importance
        parameters through BlockPrep: one kernel forward, one backward."""
        conv = self._gcn_conv()
        kc, cin = conv.weight.shape[0], conv.weight.shape[1]
        K, V = pattern.K, pattern.V
        cout = kc // K
        b = self.conv_1x1_start.weight.shape[0]
        bp = padded_bottleneck(b)
        bufs = tuple(bufs) + (None,) * (3 - len(bufs))
        imps = [i if torch.is_tensor(i) else None for i in imps] + [None] * (3 - len(imps))
        wres = bres = None
        if self._res_mode == 2:
            wres, bres = self.residual[0].weight, self.residual[0].bias
        vals, wc, biasterm, colsum, wd, bd, weff, beff, wu, wr, btr = BlockPrep.apply(
            conv.weight, conv.bias, imps[0], imps[1], imps[2], self.conv_1x1_start.weight,
            self.conv_1x1_start.bias, self.tcn_1.weight, self.tcn_1.bias, self.tcn_2.weight, self.tcn_2.bias,
            self.tcn_3.weight, self.tcn_3.bias, self.conv_1x1_end.weight, m_imp, wres, bres, bufs, pattern,
            (K, V, cin, cout, b, bp))
        w2 = conv.weight.detach().view(kc, cin)
        if cin % 32:
            w2 = F.pad(w2, (0, 32 - cin % 32))
        tc_ops = (w2, None, None) if conv.bias is None else (w2, conv.bias.detach().view(K, cout), colsum)
        if self._res_mode != 2:
            wr = btr = None
        self._prep_fused = True           # BlockPrep.backward joins the asynchronous pair-gradient kernel
        return (vals, wc, biasterm, tc_ops, wd, bd, weff, beff, wu, self.conv_1x1_end.bias, wr, btr)

    def forward_cl(self, x, adjs, m_imp, pattern, operands=None):
        """x (N*M, T, V, Cin) channels-last -> (N*M, T/stride, V, Cout)."""
        cfg = self._block_cfg(pattern)
        cfg.training = self.training
        cfg.seed = next(_seed_counter) * 0x9E3779B1 + torch.initial_seed()
        if operands is None:
            operands = self.prepare_operands(adjs, m_imp, pattern)
        cfg.pair_async = getattr(self, '_prep_fused', False)
        vals, wc, biasterm, w2, wd, bd, weff, beff, wu, bu, wr, btr = operands
        bn1, bn2 = self.tcn_start[0], self.tcn_end[0]
        bnr_w = bnr_b = None
        if self._res_mode == 2:
            rbn = self.residual[1]
            bnr_w, bnr_b = rbn.weight, rbn.bias
        out = ops.STBlock.apply(x, vals, wc, biasterm, w2, bn1.weight, bn1.bias, wd, bd, weff, beff, wu,
                                bu, bn2.weight, bn2.bias, wr, btr, bnr_w, bnr_b, cfg)
        if self.training:
            for bn in (bn1, bn2) + ((self.residual[1],) if self._res_mode == 2 else ()):
                bn.num_batches_tracked += 1
        self.last_seed = cfg.seed
        return out


def merged_temporal_taps(convs, m_imp, divisor=1.0):
    """Temporal conv weights (Cout, Cin, kt, 1) of one or several centred branches ->
    Wtt[kt_max, Cin, Cout] (tap first) and the bias.  Several branches scaled by
    ``mstcn_importance``, summed and divided by 3 (st_gcn_mstcn.py:244-247) are ONE kt_max-tap
    convolution."""
    kt = max(c.weight.shape[2] for c in convs)
    wtt, bias = 0, 0
    for i, conv in enumerate(convs):
        k = conv.weight.shape[2]
        off = (kt - k) // 2
        w = conv.weight[:, :, :, 0].permute(2, 1, 0)             # (tap, in, out)
        if off:
            w = F.pad(w, (0, 0, 0, 0, off, off))
        scale = 1 if m_imp is None else m_imp[i]
        wtt = wtt + w * scale
        bias = bias + conv.bias * scale
    if divisor != 1.0:
        wtt, bias = wtt / divisor, bias / divisor
    return wtt, bias


class FusedWideBlockMixin(object):
    """Drives ops.STBlockWide for a block with a full-width temporal convolution.  Sub-module
    names follow the reference: ``gcn``; either ``tcn`` = Sequential(BN, ReLU, Conv2d(kt x 1), BN,
    Dropout) (st_gcnold.py:160-174) or ``tcn_start`` / ``tcn_1,2,3`` / ``tcn_end``
    (st_gcn_mstcn.py:185-213); ``residual``."""

    def _init_fused(self, in_channels, out_channels, stride, dropout, residual):
        FusedBlockMixin._init_fused(self, in_channels, out_channels, stride, dropout, residual)

    _gcn_conv = FusedBlockMixin._gcn_conv

    def _bns(self):
        if hasattr(self, 'tcn'):
            return self.tcn[0], self.tcn[3]
        return self.tcn_start[0], self.tcn_end[0]

    def _block_cfg(self, pattern):
        bn1, bn2 = self._bns()
        if self._cfg is None or self._cfg.pattern is not pattern:
            bnr = ops.BNState(self.residual[1]) if self._res_mode == 2 else None
            ident = SparsePattern.identity(pattern.V, pattern.flat_idx.device)
            self._cfg = ops.BlockCfg(pattern, ident, self._io[2], self._res_mode, self._drop_p,
                                     ops.BNState(bn1), ops.BNState(bn2), bnr, 0)
        else:
            self._cfg.bn1, self._cfg.bn2 = ops.BNState(bn1), ops.BNState(bn2)
            if self._res_mode == 2:
                self._cfg.bnr = ops.BNState(self.residual[1])
        return self._cfg

    def prepare_operands(self, adjs, m_imp, pattern):
        """Kernel operands from the reference-layout parameters (side stream, see
        FusedBlockMixin.prepare_operands)."""
        conv = self._gcn_conv()
        vals, wc, biasterm, w2 = graph_conv_operands(conv.weight, conv.bias, adjs, pattern)
        if hasattr(self, 'tcn'):
            wtt, bt = merged_temporal_taps([self.tcn[2]], None)
        else:
            wtt, bt = merged_temporal_taps([self.tcn_1, self.tcn_2, self.tcn_3], m_imp,
                                           getattr(self, 'TCN_DIVISOR', 3.0))
        wr = btr = None
        if self._res_mode == 2:
            rconv = self.residual[0]
            cout, cin = rconv.weight.shape[0], rconv.weight.shape[1]
            wr = rconv.weight.view(cout, cin).t().contiguous()
            btr = rconv.bias.unsqueeze(0).expand(pattern.V, cout).contiguous()
        w2 = tuple(None if t is None else t.contiguous() for t in w2)
        return (vals.contiguous(), wc.contiguous(), biasterm.contiguous(), w2, wtt.contiguous(),
                bt.contiguous(), wr, btr)

    def forward_cl(self, x, adjs, m_imp, pattern, operands=None):
        """x (N*M, T, V, Cin) channels-last -> (N*M, T/stride, V, Cout)."""
        cfg = self._block_cfg(pattern)
        cfg.training = self.training
        cfg.seed = next(_seed_counter) * 0x9E3779B1 + torch.initial_seed()
        if operands is None:
            operands = self.prepare_operands(adjs, m_imp, pattern)
        vals, wc, biasterm, w2, wtt, bt, wr, btr = operands
        bn1, bn2 = self._bns()
        bnr_w = bnr_b = None
        if self._res_mode == 2:
            rbn = self.residual[1]
            bnr_w, bnr_b = rbn.weight, rbn.bias
        out = ops.STBlockWide.apply(x, vals, wc, biasterm, w2, bn1.weight, bn1.bias, wtt, bt,
                                    bn2.weight, bn2.bias, wr, btr, bnr_w, bnr_b, cfg)
        if self.training:
            for bn in (bn1, bn2) + ((self.residual[1],) if self._res_mode == 2 else ()):
                bn.num_batches_tracked += 1
        self.last_seed = cfg.seed
        return out


_side_streams = {}


def _side_stream(device):
    key = (device.type, device.index)
    if key not in _side_streams:
        _side_streams[key] = torch.cuda.Stream(device=device)
        # parameters are used on the side stream, so their AccumulateGrad nodes run there by
        # design; autograd joins every stream it used with the caller's stream after backward
        quiet = getattr(torch.autograd.graph, 'set_warn_on_accumulate_grad_stream_mismatch', None)
        if quiet is not None:
            quiet(False)
    return _side_streams[key]


def _side_stream_enabled():
    import os
    return os.environ.get('ISTGCN_SIDE_STREAM', '1') != '0'


def _flatten(operands):
    for t in operands:
        if isinstance(t, tuple):
            for u in _flatten(t):
                yield u
        elif t is not None:
            yield t


class FusedModelMixin(object):
    """Model.forward / extract_feature shared by every variant (st_gcnold.py:71-120): data_bn
    with the layout change, the block loop, global pooling, the ``fcn`` 1x1 conv."""

    def _pattern(self):
        dev = self.A.device
        pat = getattr(self, '_pat', None)
        key = (dev, self.A._version, self.A.data_ptr())
        if pat is None or getattr(self, '_pat_key', None) != key:
            mask = self.A != 0
            for name in ('A2', 'A3'):
                if hasattr(self, name):
                    mask = mask | (getattr(self, name) != 0)
            pat = SparsePattern(mask.cpu().numpy(), dev)
            self._pat, self._pat_key = pat, key
        return pat

    def _require_cuda(self, x):
        if not x.is_cuda:
            raise RuntimeError('istgcn_b200: this model only runs on a CUDA sm_100a device '
                               '(got a CPU tensor); there is no CPU fallback')

    def _trunk(self, x):
        self._require_cuda(x)
        N, C, T, V, M = x.size()
        x = ops.DataBN.apply(x.float(), self.data_bn.weight, self.data_bn.bias,
                             ops.BNState(self.data_bn), self.training)
        if self.training:
            self.data_bn.num_batches_tracked += 1
        pattern = self._pattern()
        n = len(self.st_gcn_networks)
        imp1 = self.edge_importance
        imp2 = getattr(self, 'edge_importance2', None)
        imp3 = getattr(self, 'edge_importance3', None)
        m_imp = getattr(self, 'mstcn_importance', None)
        def adjs_of(i):
            if hasattr(self, '_block_adjs'):        # element-power adjacency variants
                return self._block_adjs(i)
            adjs = [self.A * imp1[i]]
            if imp2 is not None:
                adjs.append(self.A2 * imp2[i])
                adjs.append(self.A3 * imp3[i])
            return adjs

        blocks = list(self.st_gcn_networks)
        fused_prep = _prep_kernel_enabled() and not hasattr(self, '_block_adjs') and m_imp is not None and \
            all(hasattr(b, 'prepare_operands_fused') for b in blocks)

        def operands_of(i, blk):
            if fused_prep:      # one kernel: raw buffers + importance parameters
                bufs = (self.A,) if imp2 is None else (self.A, self.A2, self.A3)
                imps = (imp1[i],) if imp2 is None else (imp1[i], imp2[i], imp3[i])
                return blk.prepare_operands_fused(bufs, imps, m_imp[i], pattern)
            return blk.prepare_operands(adjs_of(i), m_imp[i] if m_imp is not None else None, pattern)

        if _side_stream_enabled() and all(hasattr(b, 'prepare_operands') for b in blocks):
            # parameter regrouping of every block on a side stream, one event per block: block i
            # waits for ITS operands only, blocks i+1.. are regrouped while block i computes
            main = torch.cuda.current_stream()
            side = _side_stream(x.device)
            side.wait_stream(main)
            prepared = []
            with torch.cuda.stream(side):
                for i, blk in enumerate(blocks):
                    opnds = operands_of(i, blk)
                    for t in _flatten(opnds):
                        t.record_stream(main)
                    ev = torch.cuda.Event()
                    ev.record(side)
                    prepared.append((opnds, ev))
            for blk, (opnds, ev) in zip(blocks, prepared):
                main.wait_event(ev)
                x = blk.forward_cl(x, None, None, pattern, operands=opnds)
            return x
        for i, blk in enumerate(blocks):
            if fused_prep:
                x = blk.forward_cl(x, None, None, pattern, operands=operands_of(i, blk))
            else:
                x = blk.forward_cl(x, adjs_of(i), m_imp[i] if m_imp is not None else None, pattern)
        return x

    def forward(self, x):
        N, C, T, V, M = x.size()
        y = self._trunk(x)
        pooled = ops.Pool.apply(y, N, M)
        w = self.fcn.weight
        return F.linear(pooled, w.view(w.shape[0], w.shape[1]), self.fcn.bias)

    def extract_feature(self, x):
        N, C, T, V, M = x.size()
        y = self._trunk(x)                                   # (N*M, t, v, c)
        _, t, v, c = y.shape
        feature = y.view(N, M, t, v, c).permute(0, 4, 2, 3, 1)
        w = self.fcn.weight
        out = F.linear(y, w.view(w.shape[0], w.shape[1]), self.fcn.bias)
        output = out.view(N, M, t, v, -1).permute(0, 4, 2, 3, 1)
        return output, feature
