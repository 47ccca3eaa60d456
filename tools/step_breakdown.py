"""Per-call device time of one eager IST-GCN training step (bench.py shapes), in launch order."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'ist-gcn_b200')); sys.path.insert(0, ROOT)
import torch
import bench
import istgcn
from istgcn import _lib, trainer

dev = torch.device('cuda', 0)
istgcn.set_math('tf32')
arch = sys.argv[1] if len(sys.argv) > 1 else 'ist_gcn'
torch.manual_seed(0)
model = bench.build_model('ntu', dev, arch=arch)
tr = trainer.Trainer(model, base_lr=0.01, use_graph=False)
w = bench.WORKLOADS['ntu']
x = torch.randn(64, *w['shape'], device=dev)
y = torch.randint(0, w['num_class'], (64,), device=dev)
for _ in range(3):
    tr._iteration(x, y, True)
torch.cuda.synchronize()
order = []
orig = _lib.call
_lib.timing = {}
tr._iteration(x, y, True)
torch.cuda.synchronize()
timing, _lib.timing = _lib.timing, None
tot = 0.0
for name, evs in timing.items():
    ms = [a.elapsed_time(b) for a, b in evs]
    tot += sum(ms)
    print('%-16s n=%2d total %7.3f ms : %s' % (name, len(ms), sum(ms), ' '.join('%.3f' % m for m in ms)))
print('sum of kernels %.2f ms' % tot)
