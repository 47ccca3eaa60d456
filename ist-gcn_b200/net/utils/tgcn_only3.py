"""Drop-in for the reference ``net.utils.tgcn_only3`` (net/utils/tgcn_only3.py:76-88): the graph
convolution of tgcn.py with A**3 as the adjacency.
The powers are ELEMENT-wise, so the non-zero pattern of A is unchanged and the fused kernel runs on
the summed stack; the power is a tiny differentiable torch op in front of it (its gradient reaches
``edge_importance`` through the kernel's adjacency gradient)."""

from net.utils import tgcn as _tgcn


class ConvTemporalGraphical(_tgcn.ConvTemporalGraphical):
    def stacks(self, A, importance=None):
        a = A if importance is None else A * importance
        return [a ** 3]
