"""The reference keeps the baseline ST-GCN in net/st_gcnold.py; ``net.st_gcn`` is the same model."""
from net.st_gcn import Model, st_gcn  # noqa: F401
