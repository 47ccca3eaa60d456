"""Batch-data-parallel training, one process per GPU (replaces the reference's single-process
``nn.DataParallel`` wrap, processor/my_io.py:77-87).

The path shards over the batch axis only: every rank holds a full replica,
BatchNorm statistics stay per rank (exactly what DataParallel's per-replica BN does) and the
only exchange per iteration is the gradient average.  Gradients live in a few flat fp32
buckets filled in reverse block order; as soon as the last gradient of a bucket has been
written by the backward kernels the bucket's NCCL all-reduce is launched asynchronously, so the
collective (4.4 MB in total - latency-bound over NVLink/NVSwitch) overlaps the rest of the
backward pass.  Parameters that never receive a gradient (``gcn.branch.bn.*``, ``linear.*``:
registered by the reference but unused) are kept out of the buckets.
"""
import torch
import torch.distributed as dist

UNUSED_PARAM_MARKERS = ('.gcn.branch.bn.', '.linear.')


def is_unused(name):
    return any(m in '.' + name for m in UNUSED_PARAM_MARKERS)


class GradBuckets(object):
    """Flat gradient buckets with overlap of the all-reduce and the backward pass.

    ``named_params`` in registration order; buckets are cut in REVERSE order (the order in
    which backward produces gradients).  ``param.grad`` becomes a view into the flat buffer,
    so no gather/scatter copies are needed; call ``zero()`` instead of ``zero_grad()``.
    """

    def __init__(self, named_params, bucket_bytes=2 << 20, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        params = [(n, p) for n, p in named_params if p.requires_grad and not is_unused(n)]
        self.skipped = [n for n, p in named_params if p.requires_grad and is_unused(n)]
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
        for b_idx, b in enumerate(self.buckets):
            for n, p in b['params']:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._make_hook(b_idx)))

    def _close(self, items):
        total = sum(p.numel() for _, p in items)
        p0 = items[0][1]
        flat = torch.zeros(total, dtype=p0.dtype, device=p0.device)
        off = 0
        for _, p in items:
            p.grad = flat[off:off + p.numel()].view_as(p)
            off += p.numel()
        self.buckets.append({'flat': flat, 'params': list(items), 'pending': len(items),
                             'launched': False})

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
            self._handles.append(dist.all_reduce(b['flat'], op=dist.ReduceOp.SUM, group=self.group,
                                                 async_op=True))

    def zero(self):
        for b in self.buckets:
            b['flat'].zero_()
            b['pending'] = len(b['params'])
            b['launched'] = False
        self._handles = []

    def finish(self):
        """Call after ``backward()``: launch buckets whose parameters did not all receive a
        gradient this step, wait for the collectives and turn the sums into averages."""
        for b in self.buckets:
            if not b['launched']:
                self._launch(b)
        for h in self._handles:
            h.wait()
        self._handles = []
        if self.world > 1:
            for b in self.buckets:
                b['flat'].div_(self.world)

    def nbytes(self):
        return sum(b['flat'].numel() * b['flat'].element_size() for b in self.buckets)

    def remove(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []


def broadcast_state(module, src=0, group=None):
    """Rank ``src``'s parameters and buffers to every rank (DataParallel's ``replicate``)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src=src, group=group)
