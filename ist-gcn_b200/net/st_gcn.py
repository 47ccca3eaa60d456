"""Drop-in for the reference baseline ST-GCN (net/st_gcnold.py:13-203; net/st_gcn_twostream.py:9
imports it as ``net.st_gcn``): ST-GCN graph convolution + full-width 9x1 temporal convolution,
10 blocks.

Same constructor, sub-module names and registration order (=> same state_dict, including the
unused per-block ``linear``), same ``forward(x)`` on (N, C, T, V, M) and ``extract_feature``; every
block runs as sm_100a kernels (the temporal convolution as a sum of shifted, strided 1x1 taps)."""
import torch
import torch.nn as nn

from istgcn.modules import (FusedModelMixin, FusedWideBlockMixin, to_channels_first,
                            to_channels_last)
from net.utils.graph import Graph
from net.utils.tgcn import ConvTemporalGraphical


class Model(FusedModelMixin, nn.Module):
    r"""Model(in_channels, num_class, graph_args, edge_importance_weighting, **kwargs)
    (N, in_channels, T, V, M) -> (N, num_class).

    ``BLOCK`` / ``N_IMPORTANCE`` are the hooks of the element-power adjacency variants
    (net/st_gcn_multi3*.py, st_gcn_only3.py, st_gcn_learnA.py: the same file with another
    ``ConvTemporalGraphical``)."""
    BLOCK = None                 # set below (the class is defined after Model)
    N_IMPORTANCE = 1             # 3: edge_importance, edge_importance2, edge_importance3

    def __init__(self, in_channels, num_class, graph_args, edge_importance_weighting, **kwargs):
        super().__init__()
        st_gcn = self.BLOCK
        self.graph = Graph(**graph_args)
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
        names = ['edge_importance', 'edge_importance2', 'edge_importance3'][:self.N_IMPORTANCE]
        for name in names:
            if edge_importance_weighting:
                setattr(self, name, nn.ParameterList([
                    nn.Parameter(torch.ones(self.A.size())) for _ in self.st_gcn_networks]))
            else:
                setattr(self, name, [1] * len(self.st_gcn_networks))
        self.fcn = nn.Conv2d(256, num_class, kernel_size=1)

    def _block_adjs(self, i):
        """Adjacency stacks of block i (FusedModelMixin._trunk): the block's graph conv decides."""
        imps = [getattr(self, n)[i] for n in
                ('edge_importance', 'edge_importance2', 'edge_importance3')[:self.N_IMPORTANCE]]
        return self.st_gcn_networks[i].gcn.stacks(self.A, *imps)


class st_gcn(FusedWideBlockMixin, nn.Module):
    r"""st_gcn(in_channels, out_channels, kernel_size=(9, K), stride=1, dropout=0, residual=True);
    forward(x, A) -> (relu(x), A) on (N, C, T, V) tensors."""
    GCN = ConvTemporalGraphical

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, dropout=0, residual=True):
        super().__init__()
        assert len(kernel_size) == 2
        assert kernel_size[0] % 2 == 1
        padding = ((kernel_size[0] - 1) // 2, 0)
        self.gcn = self.GCN(in_channels, out_channels, kernel_size[1])
        self.tcn = nn.Sequential(
            nn.BatchNorm2d(out_channels),
            nn.ReLU(inplace=True),
            nn.Conv2d(out_channels, out_channels, (kernel_size[0], 1), (stride, 1), padding),
            nn.BatchNorm2d(out_channels),
            nn.Dropout(dropout, inplace=True),
        )
        self.linear = nn.Linear(3, out_channels)       # registered but never used by the reference
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

    def forward(self, x, A, *importances):
        assert A.size(0) == self.gcn.kernel_size
        pattern = self.gcn._cache.get(A)
        y = self.forward_cl(to_channels_last(x.float()), self.gcn.stacks(A, *importances), None, pattern)
        return to_channels_first(y), A


Model.BLOCK = st_gcn
