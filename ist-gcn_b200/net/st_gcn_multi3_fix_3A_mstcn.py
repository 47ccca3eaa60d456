"""Drop-in for the reference ``net.st_gcn_multi3_fix_3A_mstcn`` (net/st_gcn_multi3_fix_3A_mstcn.py):
the graph convolution of tgcn_multi3_fix_3A.py (A*importance, A**2*importance2, A**3*importance3)
+ the FULL-WIDTH Inception TCN (3 / 9 / 15 taps scaled by ``mstcn_importance`` and summed, NOT
divided by 3, :211-216), 10 blocks.  Same constructor, sub-module names and registration order
(=> same state_dict), same ``forward(x)`` on (N, C, T, V, M)."""
import torch
import torch.nn as nn

from istgcn.modules import (FusedModelMixin, FusedWideBlockMixin, to_channels_first,
                            to_channels_last)
from net.utils.graph import Graph
from net.utils.tgcn_multi3_fix_3A import ConvTemporalGraphical


class Model(FusedModelMixin, nn.Module):
    r"""Model(in_channels, num_class, graph_args, edge_importance_weighting, **kwargs)
    (N, in_channels, T, V, M) -> (N, num_class)."""

    def __init__(self, in_channels, num_class, graph_args, edge_importance_weighting, **kwargs):
        super().__init__()
        self.graph = Graph(**graph_args)
        A = torch.tensor(self.graph.A, dtype=torch.float32, requires_grad=False)
        self.register_buffer('A', A)
        kernel_size = (9, A.size(0))
        self.data_bn = nn.BatchNorm1d(in_channels * A.size(1))
        kwargs0 = {k: v for k, v in kwargs.items() if k != 'dropout'}
        table = ((64, 64, 1), (64, 64, 1), (64, 64, 1), (64, 128, 2), (128, 128, 1), (128, 128, 1),
                 (128, 256, 2), (256, 256, 1), (256, 256, 1))
        self.st_gcn_networks = nn.ModuleList(
            [st_gcn(in_channels, 64, kernel_size, 1, residual=False, **kwargs0)] +
            [st_gcn(cin, cout, kernel_size, s, **kwargs) for cin, cout, s in table])
        for name in ('edge_importance', 'edge_importance2', 'edge_importance3'):
            if edge_importance_weighting:
                setattr(self, name, nn.ParameterList([
                    nn.Parameter(torch.ones(self.A.size())) for _ in self.st_gcn_networks]))
            else:
                setattr(self, name, [1] * len(self.st_gcn_networks))
        self.mstcn_importance = nn.ParameterList([
            nn.Parameter(torch.ones(3)) for _ in self.st_gcn_networks])
        self.fcn = nn.Conv2d(256, num_class, kernel_size=1)

    def _block_adjs(self, i):
        return self.st_gcn_networks[i].gcn.stacks(self.A, self.edge_importance[i],
                                                  self.edge_importance2[i], self.edge_importance3[i])


class st_gcn(FusedWideBlockMixin, nn.Module):
    r"""st_gcn(in_channels, out_channels, kernel_size=(9, K), stride=1, dropout=0, residual=True);
    forward(x, A, importance, importance2, importance3, mstcn_importance) -> (relu(x), A)."""
    TCN_DIVISOR = 1.0

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, dropout=0, residual=True):
        super().__init__()
        assert len(kernel_size) == 2
        assert kernel_size[0] % 2 == 1
        self.gcn = ConvTemporalGraphical(in_channels, out_channels, kernel_size[1])
        self.tcn_start = nn.Sequential(nn.BatchNorm2d(out_channels), nn.ReLU(inplace=True))
        self.tcn_1 = nn.Conv2d(out_channels, out_channels, (3, 1), (stride, 1), (1, 0))
        self.tcn_2 = nn.Conv2d(out_channels, out_channels, (9, 1), (stride, 1), (4, 0))
        self.tcn_3 = nn.Conv2d(out_channels, out_channels, (15, 1), (stride, 1), (7, 0))
        self.tcn_end = nn.Sequential(nn.BatchNorm2d(out_channels), nn.Dropout(dropout, inplace=True))
        if not residual:
            self.residual = lambda x: 0
        elif (in_channels == out_channels) and (stride == 1):
            self.residual = lambda x: x
        else:
            self.residual = nn.Sequential(
                nn.Conv2d(in_channels, out_channels, kernel_size=1, stride=(stride, 1)),
                nn.BatchNorm2d(out_channels))
        self.relu = nn.ReLU(inplace=True)
        self._init_fused(in_channels, out_channels, stride, dropout, residual)

    def forward(self, x, A, importance, importance2, importance3, mstcn_importance):
        assert A.size(0) == self.gcn.kernel_size
        pattern = self.gcn._cache.get(A)
        y = self.forward_cl(to_channels_last(x.float()),
                            self.gcn.stacks(A, importance, importance2, importance3),
                            mstcn_importance, pattern)
        return to_channels_first(y), A
