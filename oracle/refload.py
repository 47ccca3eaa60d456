"""Import the REAL reference (``/root/reference``) next to this repo's own ``net`` package.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Only usable in the build container, where the
reference tree is mounted read-only; the GPU box has no /root/reference, so nothing that runs
there (``-m gpu`` tests, smoke(), bench.py) may call this.  Used by
tests/golden/make_golden.py and by the CPU tests that compare the oracle with the live
reference (skipped when the tree is absent).

Both trees own a top-level package called ``net`` and the reference uses absolute imports
(net/st_gcnold.py:6-8), so the loader swaps ``sys.modules`` around the import; the returned
module objects keep working afterwards because their classes hold their own globals.
"""
import contextlib
import importlib
import os
import sys

REF_ROOT = os.environ.get('ISTGCN_REFERENCE_ROOT', '/root/reference')
_PKGS = ('net', 'feeder', 'processor', 'torchlight', 'tools')


def available():
    return os.path.isfile(os.path.join(REF_ROOT, 'net', 'utils', 'graph.py'))


@contextlib.contextmanager
def _reference_modules():
    saved = {k: v for k, v in sys.modules.items() if k.split('.')[0] in _PKGS}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, REF_ROOT)
    old_flag = sys.dont_write_bytecode
    sys.dont_write_bytecode = True           # the reference tree is read-only
    try:
        yield
    finally:
        sys.dont_write_bytecode = old_flag
        sys.path.remove(REF_ROOT)
        for k in [k for k in sys.modules if k.split('.')[0] in _PKGS]:
            del sys.modules[k]
        sys.modules.update(saved)


def load(name):
    """load('net.st_gcnold') -> the reference module object."""
    if not available():
        raise RuntimeError('reference tree not mounted at %s' % REF_ROOT)
    with _reference_modules():
        return importlib.import_module(name)


def build_reference_model(arch, in_channels, num_class, graph_args, edge_importance_weighting=True,
                          **kwargs):
    """Instantiate the reference network for an oracle ``arch`` name.

    'ist_gcn' has no class in the reference; it is assembled from the reference's own pieces
    as SURVEY.md section 8(c) describes: the st_gcn_mstcn_1x1 block with its graph conv replaced by
    inceptionv2_gcn.Inception2, under the st_gcn_msgcn trunk plus the mstcn_importance list."""
    import torch
    import torch.nn as nn
    import torch.nn.functional as F
    table = {'st_gcn': 'net.st_gcnold', 'st_gcn_msgcn': 'net.st_gcn_msgcn',
             'st_gcn_mstcn': 'net.st_gcn_mstcn', 'st_gcn_mstcn_1x1': 'net.st_gcn_mstcn_1x1',
             'st_gcn_mstcn_1x1_deep': 'net.st_gcn_mstcn_1x1_deep',
             'st_gcn_deep_msgcn': 'net.st_gcn_deep_msgcn', 'st_gcn_msgcn_new': 'net.st_gcn_msgcn_new',
             'st_gcn_multi3': 'net.st_gcn_multi3', 'st_gcn_multi3_fix': 'net.st_gcn_multi3_fix',
             'st_gcn_only3': 'net.st_gcn_only3', 'st_gcn_learnA': 'net.st_gcn_learnA',
             'st_gcn_multi3_fix_3A': 'net.st_gcn_multi3_fix_3A',
             'st_gcn_multi3_fix_3A_mstcn': 'net.st_gcn_multi3_fix_3A_mstcn'}
    if arch in table:
        return load(table[arch]).Model(in_channels, num_class, graph_args,
                                       edge_importance_weighting, **kwargs)
    assert arch == 'ist_gcn'
    blk_mod = load('net.st_gcn_mstcn_1x1')
    inc_mod = load('net.utils.inceptionv2_gcn')
    graph_mod = load('net.utils.graph')

    class Composite(nn.Module):
        def __init__(self):
            super().__init__()
            self.graph = graph_mod.Graph(**graph_args)
            for name in ('A2', 'A3', 'A'):           # registration order of st_gcn_msgcn.py:36-41
                self.register_buffer(name, torch.tensor(getattr(self.graph, name),
                                                        dtype=torch.float32))
            K, V = self.A.size(0), self.A.size(1)
            self.data_bn = nn.BatchNorm1d(in_channels * V)
            kw0 = {k: v for k, v in kwargs.items() if k != 'dropout'}
            cfg = [(in_channels, 64, 1, False), (64, 64, 1, True), (64, 64, 1, True),
                   (64, 64, 1, True), (64, 128, 2, True), (128, 128, 1, True),
                   (128, 128, 1, True), (128, 256, 2, True), (256, 256, 1, True),
                   (256, 256, 1, True)]
            blocks = []
            for i, (ci, co, s, r) in enumerate(cfg):
                b = blk_mod.st_gcn(ci, co, (9, K), s, residual=r, **(kw0 if i == 0 else kwargs))
                b.gcn = inc_mod.Inception2(ci, co, K)
                blocks.append(b)
            self.st_gcn_networks = nn.ModuleList(blocks)
            n = len(blocks)
            if edge_importance_weighting:
                for name in ('edge_importance', 'edge_importance2', 'edge_importance3'):
                    setattr(self, name, nn.ParameterList(
                        [nn.Parameter(torch.ones(self.A.size())) for _ in range(n)]))
            else:
                self.edge_importance = self.edge_importance2 = self.edge_importance3 = [1] * n
            self.mstcn_importance = nn.ParameterList(
                [nn.Parameter(torch.ones(3)) for _ in range(n)])
            self.fcn = nn.Conv2d(256, num_class, kernel_size=1)

        def forward(self, x):
            N, C, T, V, M = x.size()
            x = x.permute(0, 4, 3, 1, 2).contiguous().view(N * M, V * C, T)
            x = self.data_bn(x)
            x = x.view(N, M, V, C, T).permute(0, 1, 3, 4, 2).contiguous().view(N * M, C, T, V)
            for blk, i1, i2, i3, m in zip(self.st_gcn_networks, self.edge_importance,
                                          self.edge_importance2, self.edge_importance3,
                                          self.mstcn_importance):
                res = blk.residual(x)
                y, _, _, _ = blk.gcn(x, self.A * i1, self.A2 * i2, self.A3 * i3)
                y = blk.conv_1x1_start(blk.tcn_start(y))
                y = blk.tcn_1(y) * m[0] + blk.tcn_2(y) * m[1] + blk.tcn_3(y) * m[2]
                y = blk.tcn_end(blk.conv_1x1_end(y))
                x = blk.relu(y + res)
            x = F.avg_pool2d(x, x.size()[2:])
            x = x.view(N, M, -1, 1, 1).mean(dim=1)
            return self.fcn(x).view(N, -1)

    return Composite()
