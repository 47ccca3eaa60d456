"""``python main.py recognition -c train.yaml`` end to end on the GPU:
the reference's command line and YAML keys, two epochs on a synthetic .npy / .pkl data set through
the feeder drop-in, the pinned prefetcher with GPU augmentation and the CUDA-graph trainer; then
resume from the checkpoint and a test-phase run.

reference: main.py:12-33, processor/processor.py:159-226, processor/my_io.py:31-87,
torchlight/torchlight/io.py:57-119."""
import os
import pickle
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MAIN = os.path.join(ROOT, 'ist-gcn_b200', 'main.py')


def _dataset(tmp_path, name, n, T=20):
    rs = np.random.RandomState(len(name))
    np.save(str(tmp_path / (name + '_data.npy')), rs.randn(n, 3, T, 25, 2).astype(np.float32))
    with open(str(tmp_path / (name + '_label.pkl')), 'wb') as f:
        pickle.dump((['%s%d' % (name, i) for i in range(n)], [int(v) for v in rs.randint(0, 60, n)]), f)


def _run(args, cwd):
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(ROOT, 'ist-gcn_b200'), ROOT]))
    r = subprocess.run([sys.executable, MAIN] + args, cwd=cwd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return r.stdout


def test_recognition_train_resume_test(tmp_path):
    from oracle import model_ref
    from net.utils.graph import Graph
    _dataset(tmp_path, 'train', 16)
    _dataset(tmp_path, 'val', 8)
    work = tmp_path / 'work'
    cfg = tmp_path / 'train.yaml'
    cfg.write_text('''work_dir: %s
feeder: feeder.feeder.Feeder
train_feeder_args:
  data_path: %s
  label_path: %s
  random_choose: True
  random_move: True
  window_size: 16
test_feeder_args:
  data_path: %s
  label_path: %s
model: net.st_gcn_mstcn_1x1.Model
model_args:
  in_channels: 3
  num_class: 60
  dropout: 0.5
  edge_importance_weighting: True
  graph_args:
    layout: 'ntu-rgb+d_sym'
    strategy: 'spatial_sym'
weight_decay: 0.0001
base_lr: 0.1
step: [1]
device: [0]
batch_size: 4
test_batch_size: 4
num_epoch: 2
save_interval: 1
eval_interval: 1
log_interval: 2
''' % (work, tmp_path / 'train_data.npy', tmp_path / 'train_label.pkl', tmp_path / 'val_data.npy',
       tmp_path / 'val_label.pkl'))
    out = _run(['recognition', '-c', str(cfg)], str(tmp_path))
    assert 'Training epoch: 0' in out and 'Training epoch: 1' in out and 'Eval epoch: 1' in out
    assert 'Top1:' in out and 'Top5:' in out and 'mean_loss' in out
    files = sorted(os.listdir(str(work)))
    assert files == ['config.yaml', 'epoch1_model.pt', 'epoch2_model.pt', 'log.txt'], files
    assert 'base_lr: 0.1' in (work / 'config.yaml').read_text()
    assert 'The model has been saved as' in (work / 'log.txt').read_text()
    # checkpoint = the reference's format: OrderedDict of CPU tensors, the reference's keys / shapes
    ckpt = torch.load(str(work / 'epoch2_model.pt'))
    g = Graph(layout='ntu-rgb+d_sym', strategy='spatial_sym')
    ref_state = model_ref.make_state('st_gcn_mstcn_1x1', 3, 60, g.A, None, None, seed=0)
    assert list(ckpt.keys()) == list(ref_state.keys())
    assert all(tuple(ckpt[k].shape) == tuple(ref_state[k].shape) and not ckpt[k].is_cuda for k in ckpt)
    assert int(ckpt['data_bn.num_batches_tracked']) == 8          # 2 epochs x 4 iterations
    first = torch.load(str(work / 'epoch1_model.pt'))
    assert not torch.equal(first['fcn.weight'], ckpt['fcn.weight'])
    # resume (config/st_gcn/ntu-xsub/train.yaml:35-36): --weights + --start_epoch, CLI over YAML
    out = _run(['recognition', '-c', str(cfg), '--weights', str(work / 'epoch2_model.pt'), '--start_epoch', '2',
                '--num_epoch', '3', '--ignore_weights', 'fcn'], str(tmp_path))
    assert 'Load weights from' in out and 'Filter [fcn] remove weights [fcn.weight].' in out
    assert 'Training epoch: 2' in out and 'Training epoch: 1' not in out
    assert os.path.isfile(str(work / 'epoch3_model.pt'))
    # test phase
    out = _run(['recognition', '-c', str(cfg), '--phase', 'test', '--weights', str(work / 'epoch3_model.pt')],
               str(tmp_path))
    assert 'Evaluation Start:' in out and 'Top1:' in out
