"""Input pipeline -> device.

reference: feeder/feeder.py:70-85 (``__getitem__``: random_choose / auto_pading / random_move in
NumPy, per sample, on the host), feeder/tools.py:32-102, processor/recognition.py:258 (a pageable,
effectively synchronous ``data.float().to(dev)`` per iteration).

Here the host only DRAWS the augmentation parameters -- with the reference's own ``random`` /
``numpy.random`` calls in the reference's order, so a seeded run consumes the two generators exactly
like ``Feeder.__getitem__`` does -- and the transform itself runs on the GPU (``istgcn_feeder_augment``:
temporal window + per-frame rotation / scale / translation of the x, y channels).  Batches travel
through two pinned staging buffers and a dedicated copy stream: the H2D copy (and the augmentation
kernel) of batch i+1 overlaps the training step of batch i.
"""
import random

import numpy as np
import torch

from ._lib import call


# ------------------------------------------------------------------------------ parameter draws
def draw_window(T, size, random_choose):
    """feeder.py:76-79 + tools.py:32-56: frame shift such that out[t] = in[t + shift] (zeros outside)
    and the output length.  Consumes ``random.randint`` exactly when the reference does."""
    if random_choose:
        if T == size:
            return 0, T
        if T < size:                                   # auto_pading(random_pad=True)
            return -random.randint(0, size - T), size
        return random.randint(0, T - size), size
    if size > 0:                                       # auto_pading(random_pad=False)
        return (0, size) if T < size else (0, T)
    return 0, T


def draw_move(T, angle_candidate=(-10., -5., 0., 5., 10.), scale_candidate=(0.9, 1.0, 1.1),
              transform_candidate=(-0.2, -0.1, 0.0, 0.1, 0.2), move_time_candidate=(1,)):
    """tools.py:59-102 up to (not including) the per-frame loop: the (T, 4) float32 table
    {cos(a)*s, sin(a)*s, t_x, t_y}; same generator calls, same order, same float64 arithmetic."""
    move_time = random.choice(list(move_time_candidate))
    node = np.arange(0, T, T * 1.0 / move_time).round().astype(int)
    node = np.append(node, T)
    num_node = len(node)
    A = np.random.choice(list(angle_candidate), num_node)
    S = np.random.choice(list(scale_candidate), num_node)
    T_x = np.random.choice(list(transform_candidate), num_node)
    T_y = np.random.choice(list(transform_candidate), num_node)
    a, s, t_x, t_y = np.zeros(T), np.zeros(T), np.zeros(T), np.zeros(T)
    for i in range(num_node - 1):
        n = node[i + 1] - node[i]
        a[node[i]:node[i + 1]] = np.linspace(A[i], A[i + 1], n) * np.pi / 180
        s[node[i]:node[i + 1]] = np.linspace(S[i], S[i + 1], n)
        t_x[node[i]:node[i + 1]] = np.linspace(T_x[i], T_x[i + 1], n)
        t_y[node[i]:node[i + 1]] = np.linspace(T_y[i], T_y[i + 1], n)
    return np.stack([np.cos(a) * s, np.sin(a) * s, t_x, t_y], axis=1).astype(np.float32)


def augment_on_device(x, shift=None, move=None, t_out=None):
    """x (N, C, Tin, V, M) fp32 on the GPU -> (N, C, t_out, V, M): window shift[n] (int32 (N,), or
    None) and per-frame move table (N, t_out, 4) fp32 (or None), on the current stream."""
    x = x.contiguous()
    N, C, Tin, V, M = x.shape
    t_out = Tin if t_out is None else int(t_out)
    out = torch.empty(N, C, t_out, V, M, device=x.device, dtype=torch.float32)
    call('feeder_augment', x, shift, move, out, N, C, Tin, t_out, V, M)
    return out


class AugmentSpec(object):
    """The feeder arguments that drive augmentation (feeder.py:36-48)."""

    def __init__(self, random_choose=False, random_move=False, window_size=-1):
        self.random_choose, self.random_move, self.window_size = bool(random_choose), bool(random_move), int(window_size)

    @property
    def active(self):
        return self.random_choose or self.random_move or self.window_size > 0

    def draw(self, n, T):
        """Parameters of a batch of n raw clips of T frames -> (shift int32 (n,), move (n, t_out, 4)
        or None, t_out), one sample after the other like the reference's DataLoader workers."""
        shifts, moves, t_out = [], [], T
        for _ in range(n):
            sh, t_out = draw_window(T, self.window_size, self.random_choose)
            shifts.append(sh)
            if self.random_move:
                moves.append(draw_move(t_out))
        shift = torch.tensor(shifts, dtype=torch.int32)
        move = torch.from_numpy(np.stack(moves)) if moves else None
        return shift, move, t_out


_copy_streams = {}


def _copy_stream(device):
    """One copy stream per device for the whole process: the caching allocator pools memory per
    stream, so a fresh stream per epoch would pay cudaMalloc for its staging tensors again."""
    key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
    if key not in _copy_streams:
        _copy_streams[key] = torch.cuda.Stream(device=device)
    return _copy_streams[key]


class DevicePrefetcher(object):
    """Iterate a host loader of (data (N, C, T, V, M), label (N,)) batches as device tensors.

    Two pinned staging slots + a copy stream: while the consumer trains on batch i, batch i+1 is
    copied into its pinned slot, sent H2D and (optionally) augmented on the copy stream.  The tensors
    of a yielded batch stay valid until the next-but-one ``next()`` (two slots).  ``h2d_bytes`` counts
    what was copied."""

    def __init__(self, loader, device, augment=None):
        self.loader, self.device = loader, torch.device(device)
        self.augment = augment if augment is not None and augment.active else None
        self.stream = _copy_stream(self.device)
        self.slots = [dict(), dict()]
        self.h2d_bytes = 0

    def _pinned(self, slot, key, like):
        buf = slot.get(key)
        if buf is None or buf.shape != like.shape or buf.dtype != like.dtype:
            buf = slot[key] = torch.empty(like.shape, dtype=like.dtype).pin_memory()
        return buf

    def _stage(self, slot, batch):
        if 'ev' in slot:
            slot['ev'].synchronize()             # the H2D copies that read this slot's pinned buffers are done
        data, label = batch
        data = torch.as_tensor(np.asarray(data) if not torch.is_tensor(data) else data)
        label = torch.as_tensor(np.asarray(label) if not torch.is_tensor(label) else label)
        data = data if data.dtype == torch.float32 else data.float()
        label = label if label.dtype == torch.int64 else label.long()
        items = {'data': data, 'label': label}
        t_out = None
        if self.augment is not None:
            shift, move, t_out = self.augment.draw(data.shape[0], data.shape[2])
            items['shift'] = shift
            if move is not None:
                items['move'] = move
        main = torch.cuda.current_stream(self.device)
        out = {}                                 # fresh device tensors (caching allocator, copy stream)
        with torch.cuda.stream(self.stream):
            for k, v in items.items():
                pin = v if v.is_pinned() else self._pinned(slot, k, v).copy_(v)
                out[k] = pin.to(self.device, non_blocking=True)
                self.h2d_bytes += v.numel() * v.element_size()
            x = out['data']
            if self.augment is not None:
                x = augment_on_device(x, out.get('shift'), out.get('move'), t_out)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        slot['ev'] = ev
        for t in list(out.values()) + [x]:
            t.record_stream(main)
        return x, out['label'], ev

    def __iter__(self):
        it = iter(self.loader)
        nxt, i = None, 0
        try:
            nxt = self._stage(self.slots[0], next(it))
        except StopIteration:
            return
        while nxt is not None:
            x, y, ev = nxt
            i += 1
            try:
                nxt = self._stage(self.slots[i % 2], next(it))     # enqueue batch i+1 before yielding i
            except StopIteration:
                nxt = None
            torch.cuda.current_stream(self.device).wait_event(ev)
            yield x, y
