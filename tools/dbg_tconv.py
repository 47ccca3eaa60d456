import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'ist-gcn_b200')); sys.path.insert(0, ROOT)
import torch, torch.nn.functional as F
from istgcn._lib import call
dev = 'cuda'
torch.manual_seed(0)
def check(NM, T, V, C, kt, s, dir_):
    Tout = (T - 1) // s + 1
    pad = (kt - 1) // 2
    w = torch.randn(C, C, kt, 1, device=dev) * 0.05          # (co, ci, kt, 1)
    b = torch.randn(C, device=dev)
    if dir_ == 1:
        x = torch.randn(NM, T, V, C, device=dev)
        ref = F.conv2d(x.permute(0, 3, 1, 2).double(), w.double(), b.double(), stride=(s, 1), padding=(pad, 0)).permute(0, 2, 3, 1)
        wrows = w[:, :, :, 0].permute(2, 0, 1).contiguous().view(kt * C, C)      # [tap][co][ci]
        out = torch.empty(NM, Tout, V, C, device=dev)
        st = torch.zeros(2, C, device=dev, dtype=torch.float64)
        call('tconv_tc', x, wrows, b, out, st[0], st[1], NM, T, Tout, V, C, C, kt, s, 1)
        torch.cuda.synchronize()
        err = ((out.double() - ref).abs().max() / ref.abs().max()).item()
        serr = ((st[0] - ref.sum((0, 1, 2))).abs().max() / ref.sum((0, 1, 2)).abs().max()).item()
        print('fwd NM %d T %d V %d C %d kt %d s %d: err %.2e stat err %.2e' % (NM, T, V, C, kt, s, err, serr))
    else:
        du = torch.randn(NM, Tout, V, C, device=dev)
        a = torch.randn(NM, T, V, C, device=dev, dtype=torch.float64, requires_grad=True)
        y = F.conv2d(a.permute(0, 3, 1, 2), w.double(), None, stride=(s, 1), padding=(pad, 0)).permute(0, 2, 3, 1)
        y.backward(du.double())
        wt = w[:, :, :, 0].permute(2, 1, 0).contiguous().view(kt * C, C)        # [tap][ci][co]
        da = torch.full((NM, T, V, C), float('nan'), device=dev)
        call('tconv_tc', du, wt, None, da, None, None, NM, T, Tout, V, C, C, kt, s, -1)
        torch.cuda.synchronize()
        err = ((da.double() - a.grad).abs().max() / a.grad.abs().max()).item()
        print('dx  NM %d T %d V %d C %d kt %d s %d: err %.2e' % (NM, T, V, C, kt, s, err))
check(3, 20, 25, 64, 9, 1, 1)
check(3, 20, 25, 64, 9, 2, 1)
check(2, 40, 25, 128, 15, 1, 1)
check(2, 10, 25, 256, 9, 1, 1)
check(3, 14, 18, 64, 9, 1, 1)
check(3, 20, 25, 64, 9, 1, -1)
check(2, 10, 25, 256, 15, 1, -1)
check(3, 20, 25, 64, 9, 2, -1)
check(2, 23, 25, 128, 15, 2, -1)
check(2, 14, 18, 64, 9, 2, -1)
def check_dw(NM, T, V, C, kt, s):
    Tout = (T - 1) // s + 1
    pad = (kt - 1) // 2
    a = torch.randn(NM, T, V, C, device=dev); du = torch.randn(NM, Tout, V, C, device=dev)
    w = torch.zeros(C, C, kt, 1, device=dev, dtype=torch.float64, requires_grad=True)
    y = F.conv2d(a.permute(0, 3, 1, 2).double(), w, None, stride=(s, 1), padding=(pad, 0)).permute(0, 2, 3, 1)
    y.backward(du.double())
    ref = w.grad[:, :, :, 0].permute(2, 1, 0).contiguous()       # [tap][ci][co]
    dW = torch.zeros(kt, C, C, device=dev); db = torch.zeros(V, C, device=dev)
    call('tconv_dw_tc', a, du, dW, db, NM, T, Tout, V, C, C, kt, s)
    torch.cuda.synchronize()
    err = ((dW.double() - ref).abs().max() / ref.abs().max()).item()
    berr = ((db.double() - du.double().sum((0, 1))).abs().max() / du.double().sum((0, 1)).abs().max()).item()
    print('dW  NM %d T %d V %d C %d kt %d s %d: err %.2e bias err %.2e' % (NM, T, V, C, kt, s, err, berr))
check_dw(3, 20, 25, 64, 9, 1)
check_dw(3, 21, 25, 64, 9, 2)
check_dw(2, 13, 25, 128, 15, 1)
check_dw(2, 10, 25, 256, 9, 1)
check_dw(3, 14, 18, 64, 9, 1)
for C, T in [(64, 300), (128, 150), (256, 75)]:
    NM, V, kt = 128, 25, 9
    a = torch.randn(NM, T, V, C, device=dev); du = torch.randn(NM, T, V, C, device=dev)
    dW = torch.zeros(kt, C, C, device=dev)
    f = lambda: call('tconv_dw_tc', a, du, dW, None, NM, T, T, V, C, C, kt, 1)
    for _ in range(3): f()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(10): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print('dW C %3d T %3d: %.3f ms  %.1f TFLOP/s' % (C, T, ms, 2.0 * NM * T * V * C * C * kt / ms / 1e9))
# timing at bench shapes
for C, T in [(64, 300), (128, 150), (256, 75)]:
    NM, V, kt = 128, 25, 9
    x = torch.randn(NM, T, V, C, device=dev); w = torch.randn(kt * C, C, device=dev) * 0.05
    out = torch.empty(NM, T, V, C, device=dev); st = torch.zeros(2, C, device=dev, dtype=torch.float64)
    f = lambda: call('tconv_tc', x, w, None, out, st[0], st[1], NM, T, T, V, C, C, kt, 1, 1)
    for _ in range(3): f()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(10): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    fl = 2.0 * NM * T * V * C * C * kt
    print('C %3d T %3d: %.3f ms  %.1f TFLOP/s  %.2f TB/s algorithmic' % (C, T, ms, fl / ms / 1e9, 2 * NM * T * V * C * 4 / ms / 1e9))
