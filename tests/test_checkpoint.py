"""Checkpoint format / class lookup of the reference driver (torchlight/torchlight/io.py:51-107,
181-189) as mirrored by istgcn.checkpoint, and the epoch driver's host logic.  CPU only."""
import os
from collections import OrderedDict

import pytest
import torch

G = dict(layout='ntu-rgb+d_sym', strategy='spatial_3_sym')


def _model():
    from istgcn import checkpoint
    return checkpoint.load_model('net.ist_gcn.Model', in_channels=3, num_class=60, graph_args=G,
                                 edge_importance_weighting=True, dropout=0.5)


def test_import_class_and_load_model():
    import net.ist_gcn
    from istgcn import checkpoint
    assert checkpoint.import_class('net.ist_gcn.Model') is net.ist_gcn.Model
    assert isinstance(_model(), net.ist_gcn.Model)
    with pytest.raises((ImportError, AttributeError)):
        checkpoint.import_class('net.no_such_module.Model')


def test_save_model_format_and_round_trip(tmp_path):
    from istgcn import checkpoint
    m = _model()
    with torch.no_grad():
        for p in m.parameters():
            p.add_(torch.randn_like(p) * 0.1)
    path = str(tmp_path / 'epoch1_model.pt')
    checkpoint.save_model(torch.nn.DataParallel(m), path)        # 'module.' prefix is stripped
    raw = torch.load(path)
    assert isinstance(raw, OrderedDict)
    assert list(raw.keys()) == list(m.state_dict().keys())
    assert all(v.device.type == 'cpu' for v in raw.values())
    m2 = _model()
    logs = []
    checkpoint.load_weights(m2, path, log=logs.append)
    for (k, a), (_, b) in zip(m.state_dict().items(), m2.state_dict().items()):
        assert torch.equal(a, b), k
    assert logs[0].startswith('Load weights from') and 'Load weights [A2].' in logs


def test_load_weights_prefix_filter_and_missing_keys(tmp_path):
    """--ignore_weights removes every tensor whose name starts with the prefix; a file that lacks
    tensors (or holds 'module.'-prefixed ones) still loads, the rest keeps the model's values."""
    from istgcn import checkpoint
    src = _model()
    with torch.no_grad():
        for p in src.parameters():
            p.add_(1.0)
    path = str(tmp_path / 'w.pt')
    sd = OrderedDict(('module.' + k, v) for k, v in src.state_dict().items() if not k.startswith('data_bn'))
    torch.save(sd, path)
    dst = _model()
    before = {k: v.clone() for k, v in dst.state_dict().items()}
    logs = []
    checkpoint.load_weights(dst, path, ignore_weights=['fcn', 'edge_importance2.0'], log=logs.append)
    after = dst.state_dict()
    for k in after:
        kept = k.startswith('fcn') or k.startswith('edge_importance2.0') or k.startswith('data_bn')
        ref = before[k] if kept else src.state_dict()[k]
        assert torch.equal(after[k], ref), k
    assert 'Filter [fcn] remove weights [fcn.weight].' in logs
    assert any(line.startswith('Can not find weights [data_bn.') for line in logs)


def test_reference_checkpoint_loads_into_reference_class(tmp_path):
    """A file written by save_model strict-loads into the REFERENCE's own class (when mounted)."""
    if not os.path.isdir('/root/reference/net'):
        pytest.skip('reference not mounted')
    from istgcn import checkpoint
    from oracle import refload
    import net.st_gcn_mstcn_1x1
    g = dict(layout='ntu-rgb+d_sym', strategy='spatial_sym')
    m = net.st_gcn_mstcn_1x1.Model(3, 60, g, True)
    path = str(tmp_path / 'epoch5_model.pt')
    checkpoint.save_model(m, path)
    ref = refload.build_reference_model('st_gcn_mstcn_1x1', 3, 60, g, True)
    ref.load_state_dict(torch.load(path), strict=True)


class _FakeBuckets(object):
    world, rank = 1, 0


def test_epoch_driver_schedule_and_files(tmp_path):
    """fit(): LR steps, save / eval cadence and file names of processor.py:170-195, on a stub
    model (the host logic does not need a GPU)."""
    from istgcn import trainer

    class Stub(trainer.Trainer):
        def __init__(self):
            self.model = torch.nn.Linear(4, 3)
            self.optimizer = torch.optim.SGD(self.model.parameters(), lr=0.1)
            self.base_lr, self.use_graph = 0.1, False
            self._graph = self._static = None
            self.buckets = _FakeBuckets()
            self.lrs = []

        def step(self, x, label):
            self.lrs.append(self.optimizer.param_groups[0]['lr'])
            return torch.nn.functional.cross_entropy(self.model(x), label)

    data = [(torch.randn(5, 4), torch.randint(0, 3, (5,))) for _ in range(3)]
    t = Stub()
    hist = t.fit(data, num_epoch=5, step=[2, 4], save_interval=2, eval_interval=3, test_loader=data,
                 work_dir=str(tmp_path), log=lambda s: None)
    assert [round(lr, 6) for lr in t.lrs[::3]] == [0.1, 0.1, 0.01, 0.01, 0.001]
    assert sorted(os.listdir(str(tmp_path))) == ['epoch2_model.pt', 'epoch4_model.pt', 'epoch5_model.pt']
    assert [('acc' in h) for h in hist] == [False, False, True, False, True]
    assert set(hist[-1]['acc']) == {1, 5} and 0.0 <= hist[-1]['acc'][1] <= 100.0
