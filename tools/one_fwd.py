import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'ist-gcn_b200')); sys.path.insert(0, ROOT)
import torch
import istgcn
from net.ist_gcn import Model
b = int(sys.argv[1]) if len(sys.argv) > 1 else 16
m = Model(3, 60, dict(layout='ntu-rgb+d_sym', strategy='spatial_3_sym'), True).cuda().train()
x = torch.randn(b, 3, 300, 25, 2, device='cuda')
for _ in range(2):
    y = m(x)
    if len(sys.argv) > 2:
        y.sum().backward()
torch.cuda.synchronize()
print('ok')
