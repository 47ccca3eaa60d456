"""Scratch timing helper (not the bench contract): time fwd / fwd+bwd of net.ist_gcn at the
BASELINE cfg-2 shape and print a per-kernel breakdown.  Usage: python tools/quick_time.py [batch]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'ist-gcn_b200'))
sys.path.insert(0, ROOT)
import torch
import torch.nn.functional as F

import istgcn
from net.ist_gcn import Model

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 64
mode = sys.argv[2] if len(sys.argv) > 2 else 'tf32'
istgcn.set_math(mode)
dev = torch.device('cuda')
torch.manual_seed(0)
model = Model(3, 60, dict(layout='ntu-rgb+d_sym', strategy='spatial_3_sym'), True, dropout=0.5).to(dev).train()
x = torch.randn(batch, 3, 300, 25, 2, device=dev)
y = torch.randint(0, 60, (batch,), device=dev)
opt = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.9, nesterov=True, weight_decay=1e-4)


def step():
    opt.zero_grad(set_to_none=True)
    loss = F.cross_entropy(model(x), y)
    loss.backward()
    opt.step()
    return loss


for _ in range(3):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
n = 5
e0.record()
for _ in range(n):
    step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print('mode %s batch %d: train step %.2f ms -> %.1f clips/s' % (mode, batch, ms, batch / ms * 1e3))
t0 = time.time()
for _ in range(n):
    step()
print('cpu launch time per step (no sync): %.2f ms' % ((time.time() - t0) / n * 1e3))
torch.cuda.synchronize()
with torch.no_grad():
    model.eval()
    for _ in range(2):
        model(x)
    e0.record()
    for _ in range(n):
        model(x)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print('eval fwd %.2f ms -> %.1f clips/s' % (ms, batch / ms * 1e3))
model.train()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(2):
        step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=30, max_name_column_width=60))
