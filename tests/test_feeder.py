"""Input pipeline and the processor's host logic (f1).

CPU: the NumPy tools and the parameter draws of istgcn.pipeline against the golden vectors the
reference's own feeder/tools.py produced (tests/golden/make_feeder_golden.py) and, when the
reference tree is mounted, against the live functions; the Feeder drop-in; the YAML <- CLI merge.
GPU (``-m gpu``): the augmentation kernel and the pinned double-buffered prefetcher against the
same golden vectors."""
import importlib.util
import os
import pickle
import random
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def _mk():
    spec = importlib.util.spec_from_file_location('make_feeder_golden',
                                                  os.path.join(HERE, 'golden', 'make_feeder_golden.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope='module')
def golden():
    return np.load(os.path.join(HERE, 'golden', 'feeder_tools.npz'))


@pytest.mark.parametrize('name', ['crop_move', 'pad_move', 'same_move', 'autopad', 'move_only', 'two_channels'])
def test_numpy_tools_match_reference_golden(name, golden):
    import feeder.tools as tools
    mk = _mk()
    out = mk.run_case(tools, name, mk.case_input(name))
    assert out.shape == golden[name].shape
    np.testing.assert_allclose(out, golden[name], rtol=0, atol=1e-6)


def test_numpy_tools_match_live_reference():
    mk = _mk()
    if not os.path.isfile(os.path.join(mk.REF, 'feeder', 'tools.py')):
        pytest.skip('reference tree not mounted')
    import feeder.tools as tools
    ref = mk.reference_tools()
    for name in mk.CASES:
        data = mk.case_input(name)
        a, b = mk.run_case(tools, name, data), mk.run_case(ref, name, data)
        np.testing.assert_allclose(a, b, rtol=0, atol=1e-6)
        # both consumed the generators identically: the next draws agree
        mk.run_case(tools, name, data)
        mine = (random.random(), np.random.rand())
        mk.run_case(ref, name, data)
        assert mine == (random.random(), np.random.rand())


def test_feeder_dropin(tmp_path):
    from feeder.feeder import Feeder
    data = np.random.RandomState(0).randn(7, 3, 20, 18, 2).astype(np.float32)
    labels = list(range(7))
    np.save(str(tmp_path / 'd.npy'), data)
    with open(str(tmp_path / 'l.pkl'), 'wb') as f:
        pickle.dump((['s%d' % i for i in labels], labels), f)
    fd = Feeder(str(tmp_path / 'd.npy'), str(tmp_path / 'l.pkl'), random_choose=True, random_move=True, window_size=12)
    assert len(fd) == 7 and (fd.N, fd.C, fd.T, fd.V, fd.M) == data.shape
    random.seed(3); np.random.seed(3)
    x, y = fd[2]
    assert x.shape == (3, 12, 18, 2) and y == 2
    raw = Feeder(str(tmp_path / 'd.npy'), str(tmp_path / 'l.pkl'), random_choose=True, random_move=True,
                 window_size=12, device_augment=True)
    xr, _ = raw[2]
    np.testing.assert_array_equal(xr, data[2])
    spec = raw.augment_spec()
    assert spec.active and spec.window_size == 12


def test_processor_argument_merge(tmp_path):
    """my_io.py:31-50: defaults <- YAML <- command line; an unknown YAML key is an assertion."""
    from processor.recognition import REC_Processor
    cfg = tmp_path / 'train.yaml'
    cfg.write_text('work_dir: ./w\nmodel: net.ist_gcn.Model\nmodel_args:\n  in_channels: 3\n  num_class: 60\n'
                   '  dropout: 0.5\nbase_lr: 0.1\nstep: [20, 40]\nbatch_size: 8\ndevice: [0]\nnum_epoch: 50\n')
    p = REC_Processor.__new__(REC_Processor)
    p.load_arg(['-c', str(cfg), '--base_lr', '0.05', '--model_args', 'dropout=0.25'])
    assert p.arg.base_lr == 0.05 and p.arg.step == [20, 40] and p.arg.batch_size == 8
    assert p.arg.model_args == {'in_channels': 3, 'num_class': 60, 'dropout': 0.25}
    assert p.arg.num_epoch == 50 and p.arg.weight_decay == 0.0001 and p.arg.nesterov is True
    bad = tmp_path / 'bad.yaml'
    bad.write_text('no_such_key: 1\n')
    with pytest.raises(AssertionError):
        REC_Processor.__new__(REC_Processor).load_arg(['-c', str(bad)])


def test_parser_has_the_reference_arguments():
    """Every ``--flag`` of the reference's three parsers (processor/my_io.py, processor.py,
    recognition.py) exists here with the same default (skipped without the reference tree)."""
    import re
    ref = os.environ.get('ISTGCN_REFERENCE_ROOT', '/root/reference')
    files = [os.path.join(ref, 'processor', f) for f in ('processor.py', 'recognition.py')]
    if not all(os.path.isfile(f) for f in files):
        pytest.skip('reference tree not mounted')
    from processor.recognition import REC_Processor
    mine = vars(REC_Processor.get_parser().parse_args([]))
    for path in files:
        for m in re.finditer(r"add_argument\((?:'-\w', )?'--(\w+)'", open(path, encoding='utf-8').read()):
            assert m.group(1) in mine, m.group(1)
    assert mine['base_lr'] == 0.01 and mine['batch_size'] == 256 and mine['show_topk'] == [1, 5]
    assert mine['feeder'] == 'feeder.feeder' and mine['log_interval'] == 100 and mine['num_epoch'] == 80


# ----------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize('name', ['crop_move', 'pad_move', 'same_move', 'autopad', 'move_only', 'two_channels'])
def test_device_augmentation_matches_reference_golden(name, golden):
    """istgcn_feeder_augment with the parameters drawn by istgcn.pipeline (same seeds) == what the
    reference's feeder/tools.py returned for that clip."""
    from istgcn import pipeline
    mk = _mk()
    shape, rc, rm, window, seed = mk.CASES[name]
    data = mk.case_input(name)
    random.seed(seed); np.random.seed(seed)
    spec = pipeline.AugmentSpec(rc, rm, window)
    shift, move, t_out = spec.draw(1, shape[1])
    x = torch.from_numpy(data)[None].cuda()
    out = pipeline.augment_on_device(x, shift.cuda(), None if move is None else move.cuda(), t_out)
    assert tuple(out.shape[1:]) == golden[name].shape
    np.testing.assert_allclose(out[0].cpu().numpy(), golden[name], rtol=0, atol=2e-6)


@pytest.mark.gpu
def test_prefetcher_double_buffer_and_augmentation():
    """DevicePrefetcher over a host loader: every batch arrives intact (pinned slots are not
    overwritten while in flight) and augmented exactly like the per-sample NumPy path with the same
    seeds; labels int64; byte count as copied."""
    import feeder.tools as tools
    from istgcn import pipeline
    rs = np.random.RandomState(1)
    batches = [(rs.randn(4, 3, 30, 25, 2).astype(np.float32), rs.randint(0, 60, 4)) for _ in range(5)]
    spec = pipeline.AugmentSpec(True, True, 16)
    random.seed(7); np.random.seed(7)
    got = []
    pf = pipeline.DevicePrefetcher(batches, 'cuda', spec)
    for x, y in pf:
        z = x * 1.0                                  # consumer work on the main stream
        got.append((z.cpu().numpy(), y.cpu().numpy()))
    assert len(got) == 5 and pf.h2d_bytes > 5 * 4 * 3 * 30 * 25 * 2 * 4
    random.seed(7); np.random.seed(7)
    for (gx, gy), (data, label) in zip(got, batches):
        assert gy.dtype == np.int64 and np.array_equal(gy, label)
        for i in range(4):
            ref = tools.random_move(np.array(tools.random_choose(np.array(data[i]), 16), dtype=np.float64))
            np.testing.assert_allclose(gx[i], ref, rtol=0, atol=2e-6)
    # no augmentation: a plain pinned double-buffered copy
    plain = [x.cpu().numpy() for x, _ in pipeline.DevicePrefetcher(batches, 'cuda')]
    for a, (data, _) in zip(plain, batches):
        np.testing.assert_array_equal(a, data)
