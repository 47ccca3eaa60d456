"""Oracle: functional plain-PyTorch restatement of the reference networks (CPU, fp32/fp64).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Everything is a pure function of a flat
``state`` dict that uses the reference's state_dict keys, so one set of
tensors can be loaded into the reference modules, this oracle and the CUDA product alike.

Reference lines followed
  graph convolution        net/utils/tgcn.py:76-89  (1x1 conv to K*C channels, k slow;
                           einsum 'nkctv,kvw->nctw')
  Inception graph conv     net/utils/inceptionv2_gcn.py:64-89 (same conv, three einsums summed;
                           the BatchNorm it owns is never applied, :30-34)
  baseline block           net/st_gcnold.py:148-203
  Inception-GCN block      net/st_gcn_msgcn.py:183-237
  Inception-TCN (full)     net/st_gcn_mstcn.py:156-249   ((x1+x2+x3)/3, :245)
  Inception-TCN (1x1)      net/st_gcn_mstcn_1x1.py:157-266 (bottleneck int(sqrt(C)), sum, :261)
  model trunk              net/st_gcnold.py:71-120, net/st_gcn_msgcn.py:99-131,
                           net/st_gcn_mstcn_1x1.py:80-106
  two-stream               net/st_gcn_twostream.py:19-26
  trainer init             processor/recognition.py:31-44
The composite 'ist_gcn' (symmetric partition + Inception GCN + 1x1 Inception TCN) does not
exist as one class in the reference (net/st_gcn_mstgcn.py is broken); it is
the composition of inceptionv2_gcn.Inception2 with the st_gcn_mstcn_1x1 block, exactly as
tests/golden/make_golden.py assembles it from the reference's own classes.
"""
from collections import OrderedDict

import torch
import torch.nn.functional as F

TEN = ((None, 64, 1, False), (64, 64, 1, True), (64, 64, 1, True), (64, 64, 1, True),
       (64, 128, 2, True), (128, 128, 1, True), (128, 128, 1, True),
       (128, 256, 2, True), (256, 256, 1, True), (256, 256, 1, True))
SEVEN = ((None, 64, 1, False), (64, 64, 1, True), (64, 64, 1, True), (64, 128, 2, True),
         (128, 128, 1, True), (128, 256, 2, True), (256, 256, 1, True))

THIRTEEN = ((None, 64, 1, False), (64, 64, 1, True), (64, 64, 1, True), (64, 64, 1, True),
            (64, 64, 1, True), (64, 128, 2, True), (128, 128, 1, True), (128, 128, 1, True),
            (128, 128, 1, True), (128, 256, 2, True), (256, 256, 1, True), (256, 256, 1, True),
            (256, 256, 1, True))

# arch -> (graph conv kind, temporal kind, block table, has the unused per-block nn.Linear)
ARCHS = {
    'st_gcn': ('single', 'plain', TEN, True),                 # net/st_gcnold.py
    'st_gcn_msgcn': ('inception', 'plain', TEN, False),       # net/st_gcn_msgcn.py
    'st_gcn_mstcn': ('single', 'incept_full', SEVEN, False),  # net/st_gcn_mstcn.py
    'st_gcn_mstcn_1x1': ('single', 'incept_1x1', TEN, False),  # net/st_gcn_mstcn_1x1.py
    'ist_gcn': ('inception', 'incept_1x1', TEN, False),       # composite, see module docstring
    # depth variants: same blocks, different block lists
    'st_gcn_mstcn_1x1_deep': ('single', 'incept_1x1', THIRTEEN, False),   # net/st_gcn_mstcn_1x1_deep.py:49-63
    'st_gcn_deep_msgcn': ('inception', 'plain', THIRTEEN, False),         # net/st_gcn_deep_msgcn.py:60-77
    'st_gcn_msgcn_new': ('inception', 'plain', SEVEN, False),             # net/st_gcn_msgcn_new.py:60-73
    # element-power adjacency variants: st_gcnold with another A_eff
    'st_gcn_multi3': ('pow_multi3', 'plain', TEN, True),           # net/utils/tgcn_multi3.py:86-89
    'st_gcn_multi3_fix': ('pow_multi3_fix', 'plain', TEN, True),   # net/utils/tgcn_multi3_fix.py:86-89
    'st_gcn_only3': ('pow_only3', 'plain', TEN, True),             # net/utils/tgcn_only3.py:86
    'st_gcn_learnA': ('pow_learnA', 'plain', TEN, True),           # net/utils/tgcn_learnA.py:75,86
    'st_gcn_multi3_fix_3A': ('pow_3A', 'plain', TEN, True),        # net/utils/tgcn_multi3_fix_3A.py:76-89
    # ... with the full-width Inception TCN, branches summed WITHOUT the /3 (:211-216)
    'st_gcn_multi3_fix_3A_mstcn': ('pow_3A', 'incept_full_sum', TEN, False),
}


def adjacency_stacks(gcn_kind, A, imps, pa=None):
    """The adjacency stacks one block aggregates with (their einsums are summed).  ``imps`` =
    (importance, importance2, importance3); powers are ELEMENT-wise (tgcn_multi3.py:87-88)."""
    a = A * imps[0]
    if gcn_kind == 'single':
        return [a]
    if gcn_kind == 'pow_multi3':
        return [a, a ** 2, a ** 3]
    if gcn_kind == 'pow_multi3_fix':       # the sum is divided by 3 afterwards (block_forward)
        return [a, a ** 2, a ** 3]
    if gcn_kind == 'pow_only3':
        return [a ** 3]
    if gcn_kind == 'pow_learnA':
        return [a ** (1 + pa)]
    if gcn_kind == 'pow_3A':
        return [a, A ** 2 * imps[1], A ** 3 * imps[2]]
    raise ValueError(gcn_kind)


def block_table(arch, in_channels):
    return [(in_channels if cin is None else cin, cout, s, res)
            for (cin, cout, s, res) in ARCHS[arch][2]]


def _bn_entries(prefix, c, gen):
    return [(prefix + 'weight', torch.empty(c).normal_(1.0, 0.02, generator=gen)),
            (prefix + 'bias', torch.zeros(c)),
            (prefix + 'running_mean', torch.zeros(c)),
            (prefix + 'running_var', torch.ones(c)),
            (prefix + 'num_batches_tracked', torch.zeros((), dtype=torch.int64))]


def _conv_entries(prefix, cout, cin, kt, gen):
    return [(prefix + 'weight', torch.empty(cout, cin, kt, 1).normal_(0.0, 0.02, generator=gen)),
            (prefix + 'bias', torch.zeros(cout))]


def make_state(arch, in_channels, num_class, A, A2=None, A3=None,
               edge_importance_weighting=True, seed=0):
    """A state_dict with the reference's keys, order and shapes, filled with
    the trainer's ``weights_init`` distribution (processor/recognition.py:31-44): Conv2d
    N(0, .02) / bias 0, BatchNorm N(1, .02) / bias 0, importances = 1.  The unused
    ``linear.*`` tensors get small random values (they never influence the output)."""
    gcn_kind, tcn_kind, _, has_linear = ARCHS[arch]
    gen = torch.Generator().manual_seed(seed)
    A = torch.as_tensor(A, dtype=torch.float32)
    K, V = A.shape[0], A.shape[1]
    items = []
    if gcn_kind == 'inception':
        items += [('A2', torch.as_tensor(A2, dtype=torch.float32)),
                  ('A3', torch.as_tensor(A3, dtype=torch.float32))]
    items.append(('A', A))
    items += _bn_entries('data_bn.', in_channels * V, gen)
    blocks = block_table(arch, in_channels)
    for i, (cin, cout, stride, residual) in enumerate(blocks):
        p = 'st_gcn_networks.%d.' % i
        if gcn_kind == 'pow_learnA':       # the module's own parameter precedes its sub-module's
            items += [(p + 'gcn.pa', torch.ones(1))]
        if gcn_kind != 'inception':
            items += _conv_entries(p + 'gcn.conv.', K * cout, cin, 1, gen)
        else:
            items += _conv_entries(p + 'gcn.branch.conv.', K * cout, cin, 1, gen)
            items += _bn_entries(p + 'gcn.branch.bn.', K * cout, gen)
        if tcn_kind == 'plain':
            items += _bn_entries(p + 'tcn.0.', cout, gen)
            items += _conv_entries(p + 'tcn.2.', cout, cout, 9, gen)
            items += _bn_entries(p + 'tcn.3.', cout, gen)
        else:
            b = int(cout ** 0.5) if tcn_kind == 'incept_1x1' else cout      # incept_full[_sum]: C -> C
            items += _bn_entries(p + 'tcn_start.0.', cout, gen)
            if tcn_kind == 'incept_1x1':
                items += _conv_entries(p + 'conv_1x1_start.', b, cout, 1, gen)
            for name, kt in (('tcn_1.', 3), ('tcn_2.', 9), ('tcn_3.', 15)):
                items += _conv_entries(p + name, b, b, kt, gen)
            if tcn_kind == 'incept_1x1':
                items += _conv_entries(p + 'conv_1x1_end.', cout, b, 1, gen)
            items += _bn_entries(p + 'tcn_end.0.', cout, gen)
        if has_linear:
            items += [(p + 'linear.weight', torch.empty(cout, 3).uniform_(-.5, .5, generator=gen)),
                      (p + 'linear.bias', torch.empty(cout).uniform_(-.5, .5, generator=gen))]
        if residual and not (cin == cout and stride == 1):
            items += _conv_entries(p + 'residual.0.', cout, cin, 1, gen)
            items += _bn_entries(p + 'residual.1.', cout, gen)
    if edge_importance_weighting:
        names = ['edge_importance'] + (['edge_importance2', 'edge_importance3']
                                       if gcn_kind in ('inception', 'pow_3A') else [])
        for name in names:
            items += [('%s.%d' % (name, i), torch.ones(K, V, V)) for i in range(len(blocks))]
    if tcn_kind != 'plain':
        items += [('mstcn_importance.%d' % i, torch.ones(3)) for i in range(len(blocks))]
    items += _conv_entries('fcn.', num_class, 256, 1, gen)
    return OrderedDict(items)


def perturb_state(state, seed=1, scale=0.2):
    """Move importances, biases and running statistics off their initial values so that parity
    tests exercise every term (a bias of exactly 0 or an importance of exactly 1 hides bugs)."""
    gen = torch.Generator().manual_seed(seed)
    out = OrderedDict()
    for k, v in state.items():
        if k in ('A', 'A2', 'A3') or k.endswith('num_batches_tracked'):
            out[k] = v.clone()
        elif 'importance' in k or k.endswith('gcn.pa'):
            out[k] = v + scale * torch.randn(v.shape, generator=gen)
        elif k.endswith('running_var'):
            out[k] = v * (1.0 + 0.5 * torch.rand(v.shape, generator=gen))
        elif k.endswith('bias') or k.endswith('running_mean'):
            out[k] = v + 0.05 * torch.randn(v.shape, generator=gen)
        else:
            out[k] = v.clone()
    return out


# ------------------------------------------------------------------------------------------
def _bn(state, prefix, x, training, momentum=0.1, eps=1e-5, update=None):
    """nn.BatchNorm{1,2}d forward (biased variance for the normalisation, unbiased for the
    running estimate).  ``update`` (a dict) receives the new running statistics."""
    w, b = state[prefix + 'weight'], state[prefix + 'bias']
    if not training:
        return F.batch_norm(x, state[prefix + 'running_mean'], state[prefix + 'running_var'],
                            w, b, False, momentum, eps)
    if update is None:
        return F.batch_norm(x, None, None, w, b, True, momentum, eps)
    rm = state[prefix + 'running_mean'].detach().clone()
    rv = state[prefix + 'running_var'].detach().clone()
    y = F.batch_norm(x, rm, rv, w, b, True, momentum, eps)
    update[prefix + 'running_mean'] = rm
    update[prefix + 'running_var'] = rv
    update[prefix + 'num_batches_tracked'] = state[prefix + 'num_batches_tracked'] + 1
    return y


def graph_conv(x, weight, bias, adjs):
    """tgcn.py:76-89 / inceptionv2_gcn.py:64-89: y = conv1x1(x) viewed (n, K, C, t, v);
    out = sum over the adjacency stacks of einsum('nkctv,kvw->nctw')."""
    K = adjs[0].shape[0]
    y = F.conv2d(x, weight, bias)
    n, kc, t, v = y.shape
    y = y.view(n, K, kc // K, t, v)
    out = None
    for A in adjs:
        term = torch.einsum('nkctv,kvw->nctw', y, A)
        out = term if out is None else out + term
    return out.contiguous()


def _dropout(x, p, training, masks, key):
    if not training or p == 0:
        return x
    if masks is not None:          # parity runs inject the keep-mask the CUDA path used
        return x * masks[key].to(x.dtype) / (1.0 - p)
    return F.dropout(x, p, True)


def block_forward(state, p, arch, x, adjs, m_imp, cfg, training, dropout=0.0, update=None,
                  masks=None):
    """One st_gcn block (see the module docstring for the per-variant reference lines)."""
    gcn_kind, tcn_kind, _, _ = ARCHS[arch]
    cin, cout, stride, residual = cfg
    if not residual:
        res = 0
    elif cin == cout and stride == 1:
        res = x
    else:
        res = F.conv2d(x, state[p + 'residual.0.weight'], state[p + 'residual.0.bias'],
                       stride=(stride, 1))
        res = _bn(state, p + 'residual.1.', res, training, update=update)
    g = 'gcn.conv.' if gcn_kind != 'inception' else 'gcn.branch.conv.'
    x = graph_conv(x, state[p + g + 'weight'], state[p + g + 'bias'], adjs)
    if gcn_kind == 'pow_multi3_fix':
        x = x / 3                                   # tgcn_multi3_fix.py:89
    if tcn_kind == 'plain':
        x = F.relu(_bn(state, p + 'tcn.0.', x, training, update=update))
        x = F.conv2d(x, state[p + 'tcn.2.weight'], state[p + 'tcn.2.bias'], stride=(stride, 1),
                     padding=(4, 0))
        x = _bn(state, p + 'tcn.3.', x, training, update=update)
    else:
        x = F.relu(_bn(state, p + 'tcn_start.0.', x, training, update=update))
        if tcn_kind == 'incept_1x1':
            x = F.conv2d(x, state[p + 'conv_1x1_start.weight'], state[p + 'conv_1x1_start.bias'])
        branches = 0
        for j, (name, pad) in enumerate((('tcn_1.', 1), ('tcn_2.', 4), ('tcn_3.', 7))):
            y = F.conv2d(x, state[p + name + 'weight'], state[p + name + 'bias'],
                         stride=(stride, 1), padding=(pad, 0))
            branches = branches + y * m_imp[j]
        if tcn_kind == 'incept_1x1':
            x = F.conv2d(branches, state[p + 'conv_1x1_end.weight'],
                         state[p + 'conv_1x1_end.bias'])
        elif tcn_kind == 'incept_full_sum':
            x = branches
        else:
            x = branches / 3
        x = _bn(state, p + 'tcn_end.0.', x, training, update=update)
    x = _dropout(x, dropout, training, masks, p)
    return F.relu(x + res)


def _trunk(state, x, arch, training, dropout, update, masks):
    """data_bn + the block loop: st_gcnold.py:74-86 (channel of data_bn = v*C + c)."""
    gcn_kind, tcn_kind, _, _ = ARCHS[arch]
    N, C, T, V, M = x.shape
    x = x.permute(0, 4, 3, 1, 2).contiguous().view(N * M, V * C, T)
    x = _bn(state, 'data_bn.', x, training, update=update)
    x = x.view(N, M, V, C, T).permute(0, 1, 3, 4, 2).contiguous().view(N * M, C, T, V)
    blocks = block_table(arch, C)
    one = torch.ones((), dtype=x.dtype)
    for i, cfg in enumerate(blocks):
        if gcn_kind.startswith('pow_'):
            imps = tuple(state.get('edge_importance%s.%d' % (sfx, i), one) for sfx in ('', '2', '3'))
            adjs = adjacency_stacks(gcn_kind, state['A'], imps,
                                    state.get('st_gcn_networks.%d.gcn.pa' % i))
        else:
            adjs = [state['A'] * state.get('edge_importance.%d' % i, one)]
        if gcn_kind == 'inception':
            adjs.append(state['A2'] * state.get('edge_importance2.%d' % i, one))
            adjs.append(state['A3'] * state.get('edge_importance3.%d' % i, one))
        m_imp = state.get('mstcn_importance.%d' % i)
        drop = 0.0 if i == 0 else dropout          # block 0 never gets the dropout kwarg
        x = block_forward(state, 'st_gcn_networks.%d.' % i, arch, x, adjs, m_imp, cfg, training,
                          drop, update, masks)
    return x


def forward(state, x, arch, training=False, dropout=0.0, update=None, masks=None):
    """Model.forward: (N, C, T, V, M) -> (N, num_class) logits (st_gcnold.py:71-96)."""
    N, _, _, _, M = x.shape
    y = _trunk(state, x, arch, training, dropout, update, masks)
    y = F.avg_pool2d(y, y.shape[2:])
    y = y.view(N, M, -1, 1, 1).mean(dim=1)
    y = F.conv2d(y, state['fcn.weight'], state['fcn.bias'])
    return y.view(y.shape[0], -1)


def extract_feature(state, x, arch):
    """Model.extract_feature (st_gcnold.py:98-120): per-(t, v, m) logits and features."""
    N, _, _, _, M = x.shape
    y = _trunk(state, x, arch, False, 0.0, None, None)
    _, c, t, v = y.shape
    feature = y.view(N, M, c, t, v).permute(0, 2, 3, 4, 1)
    out = F.conv2d(y, state['fcn.weight'], state['fcn.bias'])
    return out.view(N, M, -1, t, v).permute(0, 2, 3, 4, 1), feature


def motion_stream_input(x):
    """st_gcn_twostream.py:21-23: second-order temporal difference, zero first/last frame."""
    N, C, T, V, M = x.shape
    zero = x.new_zeros(N, C, 1, V, M)
    return torch.cat((zero, x[:, :, 1:-1] - 0.5 * x[:, :, 2:] - 0.5 * x[:, :, :-2], zero), 2)


def _sub(state, prefix):
    return OrderedDict((k[len(prefix):], v) for k, v in state.items() if k.startswith(prefix))


def twostream_forward(state, x, arch='st_gcn', training=False):
    """st_gcn_twostream.py:19-26: origin_stream(x) + motion_stream(m)."""
    return (forward(_sub(state, 'origin_stream.'), x, arch, training)
            + forward(_sub(state, 'motion_stream.'), motion_stream_input(x), arch, training))


def sgd_nesterov_step(params, grads, bufs, lr, momentum=0.9, weight_decay=1e-4):
    """optim.SGD(momentum=.9, nesterov=True, weight_decay) as configured by
    processor/recognition.py:152-159; in-place on ``params`` / ``bufs`` (None = first step)."""
    for i, (p, g) in enumerate(zip(params, grads)):
        g = g + weight_decay * p
        if bufs[i] is None:
            bufs[i] = g.clone()
        else:
            bufs[i].mul_(momentum).add_(g)
        p.sub_(lr * (g + momentum * bufs[i]))


def adjust_lr(base_lr, step, epoch):
    """processor/recognition.py:168-176."""
    return base_lr * (0.1 ** sum(1 for s in step if epoch >= s))
