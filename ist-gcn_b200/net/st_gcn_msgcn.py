"""Drop-in for the reference ``net.st_gcn_msgcn`` (net/st_gcn_msgcn.py:13-240): Inception graph
convolution (A, A^2, A^3 with three edge-importance lists) + full-width 9x1 temporal convolution,
10 blocks.  Same constructor, sub-module names and registration order (buffers A2, A3, A =>
same state_dict), same ``forward(x)`` on (N, C, T, V, M)."""
import torch
import torch.nn as nn

from istgcn.modules import (FusedModelMixin, FusedWideBlockMixin, to_channels_first,
                            to_channels_last)
from net.utils.graph import Graph
from net.utils.inceptionv2_gcn import Inception2


class Model(FusedModelMixin, nn.Module):
    r"""Model(in_channels, num_class, graph_args, edge_importance_weighting, **kwargs);
    graph_args['strategy'] must provide A2 / A3 ('spatial_3', 'spatial_3_sym')."""

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
        if edge_importance_weighting:
            self.edge_importance = nn.ParameterList([
                nn.Parameter(torch.ones(self.A.size())) for _ in self.st_gcn_networks])
            self.edge_importance2 = nn.ParameterList([
                nn.Parameter(torch.ones(self.A.size())) for _ in self.st_gcn_networks])
            self.edge_importance3 = nn.ParameterList([
                nn.Parameter(torch.ones(self.A.size())) for _ in self.st_gcn_networks])
        else:
            self.edge_importance = [1] * len(self.st_gcn_networks)
            self.edge_importance2 = [1] * len(self.st_gcn_networks)
            self.edge_importance3 = [1] * len(self.st_gcn_networks)
        self.fcn = nn.Conv2d(256, num_class, kernel_size=1)


class st_gcn(FusedWideBlockMixin, nn.Module):
    r"""st_gcn(in_channels, out_channels, kernel_size=(9, K), stride=1, dropout=0, residual=True);
    forward(x, A, A2, A3) -> (relu(x), A, A2, A3) on (N, C, T, V) tensors."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, dropout=0, residual=True):
        super().__init__()
        assert len(kernel_size) == 2
        assert kernel_size[0] % 2 == 1
        padding = ((kernel_size[0] - 1) // 2, 0)
        self.gcn = Inception2(in_channels, out_channels, kernel_size[1])
        self.tcn = nn.Sequential(
            nn.BatchNorm2d(out_channels),
            nn.ReLU(inplace=True),
            nn.Conv2d(out_channels, out_channels, (kernel_size[0], 1), (stride, 1), padding),
            nn.BatchNorm2d(out_channels),
            nn.Dropout(dropout, inplace=True),
        )
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

    def forward(self, x, A, A2, A3):
        assert A.size(0) == self.gcn.kernel_size
        pattern = self.gcn._cache.get(A, A2, A3)
        y = self.forward_cl(to_channels_last(x.float()), [A, A2, A3], None, pattern)
        return to_channels_first(y), A, A2, A3
