"""Drop-in for the reference ``net.utils.tgcn_learnA`` (net/utils/tgcn_learnA.py:75-90): the graph
convolution of tgcn.py with A**(1 + pa), pa a learnable scalar (the reference also prints pa on every call, :88; not reproduced).
The powers are ELEMENT-wise, so the non-zero pattern of A is unchanged and the fused kernel runs on
the summed stack; the power is a tiny differentiable torch op in front of it (its gradient reaches
``edge_importance`` and ``pa`` through the kernel's adjacency gradient)."""
import torch
import torch.nn as nn

from net.utils import tgcn as _tgcn


class ConvTemporalGraphical(_tgcn.ConvTemporalGraphical):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.pa = nn.Parameter(torch.ones(1))

    def stacks(self, A, importance=None):
        a = A if importance is None else A * importance
        return [a ** (1 + self.pa)]
