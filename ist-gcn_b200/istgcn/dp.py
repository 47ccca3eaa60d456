"""Batch-data-parallel training, one process per GPU (replaces the reference's single-process
``nn.DataParallel`` wrap, processor/my_io.py:77-87).

The path shards over the batch axis only: every rank holds a full replica,
BatchNorm statistics stay per rank (exactly what DataParallel's per-replica BN does) and the
only exchange per iteration is the gradient average.  Parameters and gradients live in a few
flat fp32 buckets filled in reverse block order; as soon as the last gradient of a bucket has
been written by the backward kernels the bucket's NCCL all-reduce is launched asynchronously, so
the collective (4.4 MB in total - latency-bound over NVLink/NVSwitch) overlaps the rest of the
backward pass.  The 1/world average is folded into the optimiser kernel (``FlatSGD``).
Parameters that never receive a gradient (``gcn.branch.bn.*``, ``linear.*``: registered by the
reference but unused) are kept out of the buckets.
"""
import torch
import torch.distributed as dist

UNUSED_PARAM_MARKERS = ('.gcn.branch.bn.', '.linear.')
_ALIGN = 4          # elements: every parameter starts on a 16-byte boundary of its flat buffer


def is_unused(name):
    return any(m in '.' + name for m in UNUSED_PARAM_MARKERS)


class GradBuckets(object):
    """Flat gradient (and, with ``flatten_params``, parameter) buckets with overlap of the
    all-reduce and the backward pass.

    ``named_params`` in registration order; buckets are cut in REVERSE order (the order in
    which backward produces gradients).  ``param.grad`` becomes a view into the flat buffer,
    so no gather/scatter copies are needed; call ``zero()`` instead of ``zero_grad()``.

    Host-side bucket state (pending counts, launched flags, work handles) is reset by BOTH
    ``zero()`` and ``finish()``: a CUDA-graph replay runs neither the hooks nor ``zero()`` on the
    host, so ``finish()`` must leave the state ready for the next call on its own.
    """

    def __init__(self, named_params, bucket_bytes=2 << 20, group=None, flatten_params=False):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        named_params = list(named_params)
        params = [(n, p) for n, p in named_params if p.requires_grad and not is_unused(n)]
        self.skipped = [n for n, p in named_params if p.requires_grad and is_unused(n)]
        self.flatten_params = flatten_params
        self.buckets = []                       # list of dicts: flat, params, pending
        cur, cur_bytes = [], 0
        for n, p in reversed(params):
            cur.append((n, p))
            cur_bytes += p.numel() * p.element_size()
            if cur_bytes >= bucket_bytes:
                self._close(cur)
                cur, cur_bytes = [], 0
        if cur:
            self._close(cur)
        self._handles = []
        self._hooks = []
        self.defer = False          # True: hooks only count, finish() launches every collective
        self.streams = []           # CUDA streams whose work a collective must wait for
        for b_idx, b in enumerate(self.buckets):
            for n, p in b['params']:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._make_hook(b_idx)))

    def _close(self, items):
        offs, total = [], 0
        for _, p in items:
            offs.append(total)
            total += (p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
        p0 = items[0][1]
        flat = torch.zeros(total, dtype=p0.dtype, device=p0.device)
        flat_p = None
        if self.flatten_params:
            flat_p = torch.zeros(total, dtype=p0.dtype, device=p0.device)
        for (_, p), off in zip(items, offs):
            if flat_p is not None:
                view = flat_p[off:off + p.numel()].view_as(p)
                view.copy_(p.data)
                p.data = view
            p.grad = flat[off:off + p.numel()].view_as(p)
        self.buckets.append({'flat': flat, 'flat_p': flat_p, 'params': list(items),
                             'pending': len(items), 'launched': False})

    def _make_hook(self, b_idx):
        def hook(param):
            b = self.buckets[b_idx]
            b['pending'] -= 1
            if b['pending'] == 0 and not self.defer:
                self._launch(b)
        return hook

    def _launch(self, b):
        b['launched'] = True
        if self.world > 1:
            flat = b['flat']
            if flat.is_cuda and self.streams:
                # gradients of one bucket are accumulated on more than one stream (parameter
                # regrouping and its autograd run on a side stream, FusedModelMixin._trunk): the
                # collective is ordered behind the CURRENT stream only, so join the others first
                cur = torch.cuda.current_stream(flat.device)
                for s in self.streams:
                    if s != cur:
                        cur.wait_stream(s)
            self._handles.append(dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group,
                                                 async_op=True))

    def _reset(self):
        for b in self.buckets:
            b['pending'] = len(b['params'])
            b['launched'] = False
        self._handles = []

    def zero(self):
        for b in self.buckets:
            b['flat'].zero_()
        self._reset()

    def finish(self, average=True):
        """Call after ``backward()``: launch buckets whose parameters did not all receive a
        gradient this step (or every bucket in ``defer`` mode / after a graph replay), wait for the
        collectives and, with ``average``, turn the sums into averages (``FlatSGD`` folds the
        1/world factor into its kernel instead)."""
        for b in self.buckets:
            if not b['launched']:
                self._launch(b)
        for h in self._handles:
            h.wait()
        if self.world > 1 and average:
            for b in self.buckets:
                b['flat'].div_(self.world)
        self._reset()

    def nbytes(self):
        return sum(b['flat'].numel() * b['flat'].element_size() for b in self.buckets)

    def remove(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []


class FlatSGD(object):
    """optim.SGD(momentum, nesterov, weight_decay) of processor/recognition.py:152-159 over the
    flat parameter / gradient buckets: ONE kernel per bucket (``istgcn_sgd_step``), learning rate
    in device memory (a captured graph follows the step schedule), data-parallel 1/world average
    folded in.  ``param_groups`` mirrors torch's attribute for the code that reads ``['lr']``."""

    def __init__(self, buckets, lr, momentum=0.9, nesterov=True, weight_decay=1e-4):
        assert buckets.flatten_params, 'FlatSGD needs GradBuckets(flatten_params=True)'
        self.buckets = buckets
        self.momentum, self.nesterov, self.weight_decay = float(momentum), bool(nesterov), float(weight_decay)
        dev = buckets.buckets[0]['flat'].device
        self.lr_dev = torch.full((1,), float(lr), device=dev, dtype=torch.float32)
        self.param_groups = [{'lr': float(lr), 'momentum': self.momentum, 'nesterov': self.nesterov,
                              'weight_decay': self.weight_decay,
                              'params': [p for b in buckets.buckets for _, p in b['params']]}]
        self.state = [torch.zeros_like(b['flat']) for b in buckets.buckets]     # momentum buffers

    def set_lr(self, lr):
        if self.param_groups[0]['lr'] != lr:
            self.param_groups[0]['lr'] = float(lr)
            self.lr_dev.fill_(float(lr))

    def step(self):
        from ._lib import call, i64
        lr = self.param_groups[0]['lr']
        if lr != getattr(self, '_lr_seen', lr):         # someone wrote param_groups[0]['lr'] directly
            self.lr_dev.fill_(float(lr))
        self._lr_seen = lr
        scale = 1.0 / self.buckets.world
        for b, buf in zip(self.buckets.buckets, self.state):
            call('sgd_step', b['flat_p'], b['flat'], buf, i64(b['flat'].numel()), self.lr_dev,
                 self.momentum, self.weight_decay, int(self.nesterov), scale)


def broadcast_state(module, src=0, group=None):
    """Rank ``src``'s parameters and buffers to every rank (DataParallel's ``replicate``)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src=src, group=group)


def replicas_equal(module, group=None):
    """True when every rank holds bit-identical parameters (max |p - p_rank0| == 0 on all ranks)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return True
    worst = torch.zeros(1, device=next(module.parameters()).device)
    for p in module.parameters():
        ref = p.detach().clone()
        dist.broadcast(ref, src=0, group=group)
        worst = torch.maximum(worst, (p.detach() - ref).abs().max().reshape(1))
    dist.all_reduce(worst, op=dist.ReduceOp.MAX, group=group)
    return worst.item() == 0.0
