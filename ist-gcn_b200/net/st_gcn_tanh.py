"""Drop-in for the reference ``net.st_gcn_tanh`` (net/st_gcn_tanh.py): byte-identical to
net/st_gcnold.py in the reference, so the same classes."""
from net.st_gcn import Graph, Model, st_gcn  # noqa: F401
