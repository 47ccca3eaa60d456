"""Drop-in for the reference ``net.st_gcn_twostream`` (net/st_gcn_twostream.py:11-28): two baseline
ST-GCNs, one on the joint coordinates and one on the temporal second difference ("motion"),
logits summed.  state_dict keys: ``origin_stream.*`` then ``motion_stream.*``."""
import torch
import torch.nn as nn

from .st_gcn import Model as ST_GCN


class Model(nn.Module):

    def __init__(self, *args, **kwargs):
        super().__init__()
        self.origin_stream = ST_GCN(*args, **kwargs)
        self.motion_stream = ST_GCN(*args, **kwargs)

    def forward(self, x):
        N, C, T, V, M = x.size()
        zero = x.new_zeros(N, C, 1, V, M)
        m = torch.cat((zero, x[:, :, 1:-1] - 0.5 * x[:, :, 2:] - 0.5 * x[:, :, :-2], zero), 2)
        return self.origin_stream(x) + self.motion_stream(m)
