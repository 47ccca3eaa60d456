"""Drop-in for the reference ``net.st_gcn_learnA`` (net/st_gcn_learnA.py): net/st_gcnold.py with
``net.utils.tgcn_learnA.ConvTemporalGraphical`` as the graph convolution (:6) -- same constructor,
sub-module names, registration order and ``forward(x)``; every block runs as sm_100a kernels."""
from net import st_gcn as _base
from net.utils.graph import Graph  # noqa: F401  (the reference module exposes it too)
from net.utils.tgcn_learnA import ConvTemporalGraphical


class st_gcn(_base.st_gcn):
    GCN = ConvTemporalGraphical


class Model(_base.Model):
    BLOCK = st_gcn
    N_IMPORTANCE = 1
