"""Drop-in for the reference ``net.utils.tgcn_multi3_fix_3A`` (net/utils/tgcn_multi3_fix_3A.py:76-91): the graph
convolution of tgcn.py with A*importance, A**2*importance2 and A**3*importance3 (three importance tensors, powers of the RAW A).
The powers are ELEMENT-wise, so the non-zero pattern of A is unchanged and the fused kernel runs on
the summed stack; the power is a tiny differentiable torch op in front of it (its gradient reaches
``edge_importance``, ``edge_importance2/3`` through the kernel's adjacency gradient)."""

from net.utils import tgcn as _tgcn


class ConvTemporalGraphical(_tgcn.ConvTemporalGraphical):
    def stacks(self, A, importance=None, importance2=None, importance3=None):
        one = 1 if importance is None else importance
        return [A * one, A ** 2 * (1 if importance2 is None else importance2),
                A ** 3 * (1 if importance3 is None else importance3)]
