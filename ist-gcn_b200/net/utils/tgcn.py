"""Drop-in for the reference ``net.utils.tgcn`` (net/utils/tgcn.py:7-89).

``ConvTemporalGraphical`` keeps the reference's constructor, its ``conv`` sub-module (hence the
``gcn.conv.weight`` / ``gcn.conv.bias`` state_dict keys) and its ``forward(x, A) -> (x, A)``
contract on (N, C, T, V) tensors, but the 1x1 convolution and the
``einsum('nkctv,kvw->nctw')`` (:79-86) run as ONE fused CUDA kernel.  Inside the networks the
block-level fused path (istgcn.modules.FusedBlockMixin) calls the same kernel directly in the
channels-last layout; this module-level ``forward`` is the stand-alone operator."""
import torch
import torch.nn as nn

from istgcn import ops
from istgcn.modules import graph_conv_operands, to_channels_first, to_channels_last
from istgcn.sparse import SparsePattern


class _PatternCache(object):
    """Non-zero pattern of the adjacency handed to a stand-alone graph conv."""

    def __init__(self):
        self.mask = None
        self.pattern = None

    def get(self, *adjs):
        mask = adjs[0] != 0
        for a in adjs[1:]:
            mask = mask | (a != 0)
        if self.mask is None or self.mask.shape != mask.shape or self.mask.device != mask.device \
                or not torch.equal(self.mask, mask):
            self.mask = mask
            self.pattern = SparsePattern(mask.cpu().numpy(), mask.device)
        return self.pattern


class ConvTemporalGraphical(nn.Module):
    r"""Graph convolution: 1x1 conv to K*C_out channels, then aggregation over the K-partitioned
    adjacency.  Input (N, C_in, T, V) and A (K, V, V) -> (N, C_out, T, V), A.

    Only the configuration every reference network uses is supported (t_kernel_size=1,
    t_stride=1, t_padding=0, t_dilation=1); anything else raises."""

    def __init__(self, in_channels, out_channels, kernel_size, t_kernel_size=1, t_stride=1,
                 t_padding=0, t_dilation=1, bias=True):
        super().__init__()
        if (t_kernel_size, t_stride, t_padding, t_dilation) != (1, 1, 0, 1):
            raise NotImplementedError('istgcn_b200 fuses the 1x1 graph convolution only '
                                      '(t_kernel_size=1, t_stride=1, t_padding=0, t_dilation=1)')
        self.kernel_size = kernel_size
        self.conv = nn.Conv2d(in_channels, out_channels * kernel_size, kernel_size=(t_kernel_size, 1),
                              padding=(t_padding, 0), stride=(t_stride, 1), dilation=(t_dilation, 1),
                              bias=bias)
        self._cache = _PatternCache()

    def stacks(self, A, importance=None):
        """The adjacency stacks this graph conv aggregates with (their einsums are summed; the
        element-power variants in tgcn_multi3*.py / tgcn_only3.py / tgcn_learnA.py override it)."""
        return [A if importance is None else A * importance]

    def forward(self, x, A, *importances):
        assert A.size(0) == self.kernel_size
        pattern = self._cache.get(A)
        vals, wc, biasterm, w2 = graph_conv_operands(self.conv.weight, self.conv.bias,
                                                     self.stacks(A, *importances), pattern)
        y = ops.GraphConv.apply(to_channels_last(x.float()), vals, wc, biasterm, w2, pattern)
        return to_channels_first(y), A
