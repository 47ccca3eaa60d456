"""Drop-in for the reference ``net.utils.inceptionv2_gcn`` (net/utils/inceptionv2_gcn.py:7-89).

``Inception2`` applies ONE 1x1 convolution and aggregates its output with three adjacency
stacks A, A2, A3 (graph distance 1, 2, 3), summing the results (:64-83).  Because the
aggregation is linear this equals a single aggregation with A + A2 + A3, which is how the
fused kernel runs it.  ``BasicConv2d`` keeps its never-applied BatchNorm (:30-34) so that the
``gcn.branch.bn.*`` state_dict entries exist exactly as in the reference."""
import torch.nn as nn

from istgcn import ops
from istgcn.modules import graph_conv_operands, to_channels_first, to_channels_last
from net.utils.tgcn import _PatternCache


class BasicConv2d(nn.Module):
    """conv (applied) + bn (registered, unused) -- inceptionv2_gcn.py:7-35."""

    def __init__(self, in_channels, out_channels, kernel_size, t_padding=0, t_kernel_size=1,
                 t_stride=1, t_dilation=1, bias=True):
        super().__init__()
        if (t_kernel_size, t_stride, t_padding, t_dilation) != (1, 1, 0, 1):
            raise NotImplementedError('istgcn_b200 fuses the 1x1 graph convolution only')
        self.conv = nn.Conv2d(in_channels, out_channels * kernel_size, kernel_size=(t_kernel_size, 1),
                              padding=(t_padding, 0), stride=(t_stride, 1), dilation=(t_dilation, 1),
                              bias=bias)
        self.bn = nn.BatchNorm2d(out_channels * kernel_size)

    def forward(self, x):
        raise RuntimeError('BasicConv2d is only evaluated fused inside Inception2 (the K*C_out '
                           'intermediate is never materialised)')


class Inception2(nn.Module):
    """forward(x, A, A2, A3) -> (out, A, A2, A3) on (N, C, T, V) tensors."""

    def __init__(self, in_channels, out_channels, kernel_size, t_padding=0, t_kernel_size=1,
                 t_stride=1, t_dilation=1, bias=True):
        super().__init__()
        self.kernel_size = kernel_size
        self.branch = BasicConv2d(in_channels, out_channels, kernel_size)
        self._cache = _PatternCache()

    def forward(self, x, A, A2, A3):
        assert A.size(0) == self.kernel_size
        pattern = self._cache.get(A, A2, A3)
        conv = self.branch.conv
        vals, wc, biasterm, w2 = graph_conv_operands(conv.weight, conv.bias, [A, A2, A3], pattern)
        y = ops.GraphConv.apply(to_channels_last(x.float()), vals, wc, biasterm, w2, pattern)
        return to_channels_first(y), A, A2, A3
