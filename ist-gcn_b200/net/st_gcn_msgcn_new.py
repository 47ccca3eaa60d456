"""Drop-in for the reference ``net.st_gcn_msgcn_new`` (net/st_gcn_msgcn_new.py:13-246): the 7-block
variant of ``net.st_gcn_msgcn``; its ``inceptionv2_gcn_new.Inception2`` computes the same three
einsums (net/utils/inceptionv2_gcn_new.py:46-57)."""
import torch
import torch.nn as nn

from istgcn.modules import FusedModelMixin
from net.utils.graph import Graph
from net.st_gcn_msgcn import st_gcn


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
            st_gcn(64, 128, kernel_size, 2, **kwargs),
            st_gcn(128, 128, kernel_size, 1, **kwargs),
            st_gcn(128, 256, kernel_size, 2, **kwargs),
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
