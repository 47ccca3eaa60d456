"""Checkpoint format and class lookup of the reference's driver, so that ``epochN_model.pt`` files
move both ways between the reference and these drop-in models.

reference: torchlight/torchlight/io.py:51-107 (load_model, load_weights, save_model) and
:181-189 (import_class)."""
import importlib
from collections import OrderedDict

import torch


def import_class(name):
    """'net.ist_gcn.Model' -> the class (io.py:181-189: the first component is imported, the rest
    are attribute look-ups)."""
    components = name.split('.')
    mod = importlib.import_module('.'.join(components[:-1]))
    return getattr(mod, components[-1])


def load_model(model, **model_args):
    """io.py:51-55: instantiate the class named by the YAML ``model`` key with ``model_args``."""
    return import_class(model)(**model_args)


def save_model(model, path):
    """io.py:100-106: an OrderedDict name -> CPU tensor with every 'module.' removed (the
    DataParallel prefix), written with ``torch.save``."""
    weights = OrderedDict([[''.join(k.split('module.')), v.cpu()] for k, v in model.state_dict().items()])
    torch.save(weights, path)
    return weights


def load_weights(model, weights_path, ignore_weights=None, log=None):
    """io.py:57-90: strip the 'module.' prefix, drop every tensor whose name STARTS with one of
    ``ignore_weights``, then load; tensors the file lacks (or that were filtered) keep the model's
    own values.  ``log`` (callable) receives the reference's log lines."""
    log = log or (lambda s: None)
    if ignore_weights is None:
        ignore_weights = []
    if isinstance(ignore_weights, str):
        ignore_weights = [ignore_weights]
    log('Load weights from {}.'.format(weights_path))
    weights = torch.load(weights_path, map_location='cpu')
    weights = OrderedDict([[k.split('module.')[-1], v.cpu()] for k, v in weights.items()])
    for i in ignore_weights:
        for n in [w for w in weights if w.find(i) == 0]:
            weights.pop(n)
            log('Filter [{}] remove weights [{}].'.format(i, n))
    for w in weights:
        log('Load weights [{}].'.format(w))
    try:
        model.load_state_dict(weights)
    except (KeyError, RuntimeError):
        state = model.state_dict()
        for d in set(state.keys()).difference(set(weights.keys())):
            log('Can not find weights [{}].'.format(d))
        state.update(weights)
        model.load_state_dict(state)
    return model
