"""Generate the golden fixtures in this directory FROM THE REFERENCE ITSELF.

Run in the build container (the only place /root/reference exists):

    python tests/golden/make_golden.py

Writes
  graphs.npz        every working Graph(layout, strategy): A / A2 / A3 float64 arrays, produced
                    by the reference's net.utils.graph.Graph
  model_<arch>.npz  for each network: the seeded input, the logits (train and eval mode), the
                    loss, the updated data_bn running statistics and, per parameter, gradient
                    probes (L2 norm, sum, 48 strided samples) -- produced by the reference's own
                    nn.Modules loaded (strict) with an oracle-generated state_dict.  Weights are
                    NOT stored (4 MB per net); the fixture records the SHA-256 of the state the
                    generator used and tests regenerate it from the seed and compare the hash.
The reference ships no fixtures of its own, so these files are what pins the
oracle (oracle/graph_ref.py, oracle/model_ref.py) to the reference's behaviour.
"""
import hashlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import model_ref, refload  # noqa: E402

GRAPH_CASES = [(l, s) for l in ('openpose', 'openpose_sym', 'ntu-rgb+d', 'ntu-rgb+d_sym',
                                'ntu-rgb+d_half')
               for s in ('uniform', 'distance', 'spatial', 'spatial_half', 'spatial_3',
                         'spatial_sym', 'spatial_3_sym')]

# arch -> (graph_args, num_class, input shape (N, C, T, V, M))
MODEL_CASES = {
    'st_gcn': (dict(layout='ntu-rgb+d', strategy='spatial'), 60, (2, 3, 24, 25, 2)),
    'st_gcn_msgcn': (dict(layout='ntu-rgb+d_sym', strategy='spatial_3_sym'), 60, (2, 3, 24, 25, 2)),
    'st_gcn_mstcn': (dict(layout='ntu-rgb+d', strategy='spatial'), 60, (2, 3, 24, 25, 2)),
    'st_gcn_mstcn_1x1': (dict(layout='ntu-rgb+d_sym', strategy='spatial_sym'), 60,
                         (2, 3, 24, 25, 2)),
    'ist_gcn': (dict(layout='ntu-rgb+d_sym', strategy='spatial_3_sym'), 60, (2, 3, 24, 25, 2)),
    'ist_gcn_kinetics': (dict(layout='openpose_sym', strategy='spatial_3_sym'), 400,
                         (2, 3, 20, 18, 2)),
    'st_gcn_mstcn_1x1_deep': (dict(layout='ntu-rgb+d_sym', strategy='spatial_sym'), 60,
                              (2, 3, 16, 25, 2)),
    'st_gcn_deep_msgcn': (dict(layout='ntu-rgb+d_sym', strategy='spatial_3_sym'), 60, (2, 3, 16, 25, 2)),
    'st_gcn_msgcn_new': (dict(layout='openpose_sym', strategy='spatial_3_sym'), 60, (2, 3, 16, 18, 2)),
    'st_gcn_multi3': (dict(layout='ntu-rgb+d', strategy='spatial'), 60, (2, 3, 16, 25, 2)),
    'st_gcn_multi3_fix': (dict(layout='openpose', strategy='spatial'), 60, (2, 3, 16, 18, 2)),
    'st_gcn_only3': (dict(layout='ntu-rgb+d', strategy='spatial'), 60, (2, 3, 16, 25, 2)),
    'st_gcn_learnA': (dict(layout='ntu-rgb+d', strategy='spatial'), 60, (2, 3, 16, 25, 2)),
    'st_gcn_multi3_fix_3A': (dict(layout='ntu-rgb+d_sym', strategy='spatial_sym'), 60, (2, 3, 16, 25, 2)),
    'st_gcn_multi3_fix_3A_mstcn': (dict(layout='openpose', strategy='spatial'), 60, (2, 3, 16, 18, 2)),
}


def state_digest(state):
    h = hashlib.sha256()
    for k, v in state.items():
        h.update(k.encode())
        h.update(np.ascontiguousarray(v.detach().cpu().numpy()).tobytes())
    return h.hexdigest()


def probe(t):
    flat = t.detach().reshape(-1).double()
    idx = torch.linspace(0, flat.numel() - 1, min(48, flat.numel())).long()
    return np.concatenate([[flat.norm().item(), flat.sum().item()], flat[idx].numpy()])


def case_inputs(name, shape, num_class):
    gen = torch.Generator().manual_seed(1234 + sum(map(ord, name)))
    x = torch.randn(shape, generator=gen)
    label = torch.randint(0, num_class, (shape[0],), generator=gen)
    return x, label


def case_state(name, graph):
    arch = name.replace('_kinetics', '')
    g_args, num_class, shape = MODEL_CASES[name]
    st = model_ref.make_state(arch, shape[1], num_class, graph.A, getattr(graph, 'A2', None),
                              getattr(graph, 'A3', None), seed=7)
    return model_ref.perturb_state(st, seed=11)


def main():
    torch.manual_seed(0)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    graph_mod = refload.load('net.utils.graph')
    only = set(sys.argv[1:])            # optional: regenerate only the named model cases
    out = {}
    for layout, strategy in ([] if only else GRAPH_CASES):
        g = graph_mod.Graph(layout=layout, strategy=strategy)
        for name in ('A', 'A2', 'A3'):
            if hasattr(g, name):
                out['%s|%s|%s' % (layout, strategy, name)] = np.ascontiguousarray(getattr(g, name))
    if not only:
        np.savez_compressed(os.path.join(HERE, 'graphs.npz'), **out)
        print('graphs.npz', len(out), 'arrays')

    for name, (g_args, num_class, shape) in MODEL_CASES.items():
        if only and name not in only:
            continue
        arch = name.replace('_kinetics', '')
        graph = graph_mod.Graph(**g_args)
        state = case_state(name, graph)
        x, label = case_inputs(name, shape, num_class)
        model = refload.build_reference_model(arch, shape[1], num_class, g_args, True)
        missing = model.load_state_dict(state, strict=True)
        assert list(model.state_dict().keys()) == list(state.keys()), 'key order differs'
        fix = {'x': x.numpy(), 'label': label.numpy(), 'state_sha256': np.array(state_digest(state))}
        model.eval()
        with torch.no_grad():
            fix['logits_eval'] = model(x).numpy()
            if arch == 'st_gcn':                 # only st_gcnold has a working extract_feature
                o, f = model.extract_feature(x)
                fix['feat_out'] = probe(o)
                fix['feat_feature'] = probe(f)
        model.train()
        logits = model(x)
        loss = torch.nn.functional.cross_entropy(logits, label)
        loss.backward()
        fix['logits_train'] = logits.detach().numpy()
        fix['loss'] = np.array(loss.item())
        fix['data_bn.running_mean'] = model.data_bn.running_mean.numpy().copy()
        fix['data_bn.running_var'] = model.data_bn.running_var.numpy().copy()
        sd_after = model.state_dict()
        fix['block0_bn_running_var'] = [v for k, v in sd_after.items()
                                        if k.startswith('st_gcn_networks.0.') and
                                        k.endswith('running_var') and 'branch' not in k][0].numpy().copy()
        names = []
        for k, p in model.named_parameters():
            if p.grad is None:
                continue
            names.append(k)
            fix['grad|' + k] = probe(p.grad)
        fix['grad_names'] = np.array(names)
        np.savez_compressed(os.path.join(HERE, 'model_%s.npz' % name), **fix)
        print(name, 'loss %.6f' % loss.item(), len(names), 'grads',
              os.path.getsize(os.path.join(HERE, 'model_%s.npz' % name)) // 1024, 'KB')


if __name__ == '__main__':
    main()
