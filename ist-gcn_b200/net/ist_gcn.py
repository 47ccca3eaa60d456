"""``net.ist_gcn`` -- the full IST-GCN: symmetric partition + Inception GCN + Inception TCN.

The reference intends this network in net/st_gcn_mstgcn.py, but that file cannot run
(``self.tcn(x)`` :237 vs ``MSTCN.forward(x, mstcn_importance)`` net/utils/ms_tcn.py:41;
SURVEY.md section 0.4).  This module is the composition the README describes, assembled from the
reference's working pieces: the trunk and the three ``edge_importance*`` lists of
net/st_gcn_msgcn.py:29-131, ``Inception2`` (net/utils/inceptionv2_gcn.py:38-89) as the graph
convolution and the 1x1-bottleneck Inception-TCN block of net/st_gcn_mstcn_1x1.py:157-266 with
its ``mstcn_importance`` list (:71-74).  state_dict = the union of those layouts
(SURVEY.md App. B): buffers registered in the order A2, A3, A."""
import torch
import torch.nn as nn

from istgcn.modules import (FusedBlockMixin, FusedModelMixin, to_channels_first,
                            to_channels_last)
from net.utils.graph import Graph
from net.utils.inceptionv2_gcn import Inception2


class Model(FusedModelMixin, nn.Module):
    r"""Model(in_channels, num_class, graph_args, edge_importance_weighting, **kwargs);
    graph_args['strategy'] must be 'spatial_3' or 'spatial_3_sym' (A2 / A3 are required)."""

    def __init__(self, in_channels, num_class, graph_args, edge_importance_weighting, **kwargs):
        super().__init__()
        self.graph = Graph(**graph_args)
        A2 = torch.tensor(self.graph.A2, dtype=torch.float32, requires_grad=False)
        A3 = torch.tensor(self.graph.A3, dtype=torch.float32, requires_grad=False)
        self.register_buffer('A2', A2)
        self.register_buffer('A3', A3)
        A = torch.tensor(self.graph.A, dtype=torch.float32, requires_grad=False)
        self.register_buffer('A', A)
        spatial_kernel_size = A.size(0)
        temporal_kernel_size = 9
        kernel_size = (temporal_kernel_size, spatial_kernel_size)
        self.data_bn = nn.BatchNorm1d(in_channels * A.size(1))
        kwargs0 = {k: v for k, v in kwargs.items() if k != 'dropout'}
        self.st_gcn_networks = nn.ModuleList((
            st_gcn(in_channels, 64, kernel_size, 1, residual=False, **kwargs0),
            st_gcn(64, 64, kernel_size, 1, **kwargs),
            st_gcn(64, 64, kernel_size, 1, **kwargs),
            st_gcn(64, 64, kernel_size, 1, **kwargs),
            st_gcn(64, 128, kernel_size, 2, **kwargs),
            st_gcn(128, 128, kernel_size, 1, **kwargs),
            st_gcn(128, 128, kernel_size, 1, **kwargs),
            st_gcn(128, 256, kernel_size, 2, **kwargs),
            st_gcn(256, 256, kernel_size, 1, **kwargs),
            st_gcn(256, 256, kernel_size, 1, **kwargs),
        ))
        n = len(self.st_gcn_networks)
        if edge_importance_weighting:
            self.edge_importance = nn.ParameterList(
                [nn.Parameter(torch.ones(self.A.size())) for _ in range(n)])
            self.edge_importance2 = nn.ParameterList(
                [nn.Parameter(torch.ones(self.A2.size())) for _ in range(n)])
            self.edge_importance3 = nn.ParameterList(
                [nn.Parameter(torch.ones(self.A3.size())) for _ in range(n)])
        else:
            self.edge_importance = [1] * n
            self.edge_importance2 = [1] * n
            self.edge_importance3 = [1] * n
        self.mstcn_importance = nn.ParameterList([nn.Parameter(torch.ones(3)) for _ in range(n)])
        self.fcn = nn.Conv2d(256, num_class, kernel_size=1)


class st_gcn(FusedBlockMixin, nn.Module):
    r"""Inception2 graph conv + 1x1-bottleneck Inception TCN;
    forward(x, A, A2, A3, mstcn_importance) -> (relu(x), A, A2, A3) on (N, C, T, V) tensors."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, dropout=0, residual=True):
        super().__init__()
        assert len(kernel_size) == 2
        assert kernel_size[0] % 2 == 1
        b = int(out_channels ** 0.5)
        self.gcn = Inception2(in_channels, out_channels, kernel_size[1])
        self.tcn_start = nn.Sequential(nn.BatchNorm2d(out_channels), nn.ReLU(inplace=True))
        self.conv_1x1_start = nn.Conv2d(out_channels, b, (1, 1), (1, 1), (0, 0))
        self.tcn_1 = nn.Conv2d(b, b, (3, 1), (stride, 1), (1, 0))
        self.tcn_2 = nn.Conv2d(b, b, (9, 1), (stride, 1), (4, 0))
        self.tcn_3 = nn.Conv2d(b, b, (15, 1), (stride, 1), (7, 0))
        self.conv_1x1_end = nn.Conv2d(b, out_channels, (1, 1), (1, 1), (0, 0))
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

    def forward(self, x, A, A2, A3, mstcn_importance):
        assert A.size(0) == self.gcn.kernel_size
        pattern = self.gcn._cache.get(A, A2, A3)
        y = self.forward_cl(to_channels_last(x.float()), [A, A2, A3], mstcn_importance, pattern)
        return to_channels_first(y), A, A2, A3
