"""The training hot loop of the reference's REC_Processor (processor/recognition.py:185-310),
restated for one process per GPU: same initialisation (``weights_init`` :31-44), loss
(CrossEntropy :150), optimiser (SGD momentum 0.9, nesterov, weight decay :152-159) and step
schedule (``adjust_lr`` :168-176); the DataParallel wrap is replaced by ``dp.GradBuckets``.
Logging / TensorBoard / confusion matrices of the reference are out of scope.
"""
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import dp


def weights_init(m):
    """processor/recognition.py:31-44: Conv1d / (exactly) Conv2d weights ~ N(0, .02), bias 0;
    every *BatchNorm* weight ~ N(1, .02), bias 0."""
    classname = m.__class__.__name__
    if classname.find('Conv1d') != -1 or type(m) == nn.Conv2d:
        m.weight.data.normal_(0.0, 0.02)
        if m.bias is not None:
            m.bias.data.fill_(0)
    elif classname.find('BatchNorm') != -1:
        m.weight.data.normal_(1.0, 0.02)
        m.bias.data.fill_(0)


def adjust_lr(optimizer, base_lr, step, epoch):
    """lr = base_lr * 0.1 ** #{s in step : epoch >= s}  (recognition.py:168-176)."""
    lr = base_lr * (0.1 ** sum(1 for s in (step or []) if epoch >= s))
    if hasattr(optimizer, 'set_lr'):          # dp.FlatSGD: the rate lives in device memory
        optimizer.set_lr(lr)
    else:
        for group in optimizer.param_groups:
            group['lr'] = lr
    return lr


class Trainer(object):
    """model + SGD(nesterov) + gradient buckets; ``step(x, label)`` is one iteration of
    recognition.py:249-298 (forward, loss, zero_grad, backward, optimizer step).

    ``use_graph=True`` captures the iteration in one CUDA graph per input shape after two eager
    warm-up steps at that shape and replays it afterwards: a step is ~200 kernels plus the
    parameter regrouping, which is launch-bound from Python at B200 speeds.  The learning rate is
    read from device memory by the optimiser kernel, so the step schedule needs no re-capture.
    One rank: the graph holds forward, backward and the optimiser kernel.  Several ranks: the
    graph holds forward + backward; the bucket all-reduces (4.4 MB, latency-bound) and the
    optimiser kernel run right after the replay.  ``ISTGCN_GRAPH_COLLECTIVES=1`` captures the
    collectives as well (launched from the gradient hooks on NCCL's stream, overlapping the rest
    of the backward pass inside the graph); a graph that holds NCCL kernels must be destroyed
    BEFORE the process group -- call ``close()`` -- or ``destroy_process_group`` blocks forever
    (observed on this pool, torch 2.11 / NCCL 2.28.9)."""

    def __init__(self, model, base_lr=0.1, weight_decay=1e-4, nesterov=True, momentum=0.9,
                 bucket_bytes=2 << 20, group=None, use_graph=False):
        self.model = model
        self.buckets = dp.GradBuckets(list(model.named_parameters()), bucket_bytes, group,
                                      flatten_params=True)
        self.optimizer = dp.FlatSGD(self.buckets, base_lr, momentum=momentum, nesterov=nesterov,
                                    weight_decay=weight_decay)
        self.base_lr = base_lr
        self.use_graph = use_graph
        self._graphs = {}           # (x.shape, label.shape) -> (graph, static x, static label, loss)
        self._eager_seen = {}
        self.graph_collectives = os.environ.get('ISTGCN_GRAPH_COLLECTIVES', '0') == '1'

    # one iteration; ``with_optimizer`` False leaves the summed gradients in the buckets
    def _iteration(self, x, label, with_optimizer=True):
        from . import modules, ops
        self.model.train()
        if x.is_cuda:
            streams = [torch.cuda.current_stream(x.device)]
            if modules._side_stream_enabled():
                streams.append(modules._side_stream(x.device))
            self.buckets.streams = streams
        ops.step_counter(x.device).add_(1)
        self.buckets.zero()
        output = self.model(x)
        loss = F.cross_entropy(output, label)
        loss.backward()
        if with_optimizer:
            self.buckets.finish(average=False)      # FlatSGD applies the 1/world factor
            self.optimizer.step()
        return loss

    def invalidate_graph(self):
        """Drop every captured graph (e.g. after the model's structure or mode flags changed)."""
        self._graphs = {}

    def close(self):
        """Release the captured graphs (required before ``destroy_process_group`` when they hold
        NCCL collectives) and the gradient hooks."""
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        self._graphs = {}
        self._eval_graph = None
        self.buckets.remove()
        import gc
        gc.collect()
        if torch.cuda.is_available():
            torch.cuda.synchronize()

    def set_lr(self, lr):
        if hasattr(self.optimizer, 'set_lr'):
            self.optimizer.set_lr(lr)
        else:
            for group in self.optimizer.param_groups:
                group['lr'] = lr

    def step(self, x, label):
        if not self.use_graph:
            return self._iteration(x, label)
        key = (tuple(x.shape), tuple(label.shape))
        ent = self._graphs.get(key)
        if ent is None:
            seen = self._eager_seen.get(key, 0)
            if seen < 2:                # momentum buffers, caches, cuda handles, lazy attributes
                self._eager_seen[key] = seen + 1
                return self._iteration(x, label)
            whole = self.buckets.world == 1 or self.graph_collectives
            sx, sy = torch.empty_like(x), torch.empty_like(label)
            sx.copy_(x)
            sy.copy_(label)
            self.buckets.defer = not whole
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                sloss = self._iteration(sx, sy, whole)
            ent = self._graphs[key] = (graph, sx, sy, sloss, whole)
            # the capture itself did not execute anything: fall through to the first replay
        graph, sx, sy, sloss, whole = ent
        sx.copy_(x, non_blocking=True)
        sy.copy_(label, non_blocking=True)
        graph.replay()
        if not whole:
            self.buckets.finish(average=False)
            self.optimizer.step()
        return sloss

    @torch.no_grad()
    def evaluate(self, x):
        """Inference forward.  With ``use_graph`` the eval forward of this input shape is captured
        once and replayed (it is launch-bound from Python otherwise: ~190 kernels + the parameter
        regrouping per call).  Parameters and running statistics are updated in place by the
        training step, so a replay always reads their current values; a new input shape
        re-captures."""
        self.model.eval()
        if not self.use_graph:
            return self.model(x)
        key = (tuple(x.shape), x.dtype, x.device)
        ev = getattr(self, '_eval_graph', None)
        if ev is None or ev[0] != key:
            for _ in range(2):                      # lazy initialisations outside the capture
                self.model(x)
            sx = torch.empty_like(x)
            sx.copy_(x)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                out = self.model(sx)
            ev = self._eval_graph = (key, graph, sx, out)
        _, graph, sx, out = ev
        sx.copy_(x, non_blocking=True)
        graph.replay()
        return out

    # ------------------------------------------------------------------ epoch driver
    def train_epoch(self, loader, epoch=0, step=None, device=None, augment=None):
        """One pass of recognition.py:185-310 over ``loader`` (yields host (data, label) batches):
        step LR schedule, pinned double-buffered H2D on a copy stream (istgcn.pipeline), one ``step``
        per batch.  The loss is read back once per epoch, not once per iteration (the reference's
        per-iteration ``.item()`` is a host sync).  Returns the mean loss."""
        device = device or next(self.model.parameters()).device
        self.set_lr(self.base_lr * (0.1 ** sum(1 for s in (step or []) if epoch >= s)))
        total, count = None, 0
        if torch.device(device).type == 'cuda':
            from . import pipeline
            batches = pipeline.DevicePrefetcher(loader, device, augment)
        else:
            batches = ((d.float().to(device), l.long().to(device)) for d, l in loader)
        for data, label in batches:
            loss = self.step(data, label).detach()
            total = loss.clone() if total is None else total + loss
            count += 1
        return float(total.item()) / max(count, 1) if count else float('nan')

    @torch.no_grad()
    def test_epoch(self, loader, topk=(1, 5), device=None):
        """recognition.py:312-345 + show_topk (:178-183): mean loss and top-k accuracies (%)."""
        device = device or next(self.model.parameters()).device
        self.model.eval()
        results, labels, losses = [], [], []
        for data, label in loader:
            data = data.float().to(device, non_blocking=True)
            label = label.long().to(device, non_blocking=True)
            out = self.model(data)
            losses.append(F.cross_entropy(out, label))
            results.append(out)
            labels.append(label)
        result, label = torch.cat(results), torch.cat(labels)
        rank = result.argsort(dim=1)
        acc = {k: 100.0 * (rank[:, -k:] == label[:, None]).any(dim=1).float().mean().item() for k in topk}
        return torch.stack(losses).mean().item(), acc

    def fit(self, train_loader, num_epoch, step=None, start_epoch=0, save_interval=10, eval_interval=5,
            test_loader=None, work_dir=None, log=print):
        """processor/processor.py:159-226 (train phase): per epoch train, save
        ``epoch{N}_model.pt`` every ``save_interval`` epochs and at the end, evaluate every
        ``eval_interval`` epochs and at the end.  Rank 0 writes the files."""
        import os
        from . import checkpoint
        history = []
        for epoch in range(start_epoch, num_epoch):
            log('Training epoch: {}'.format(epoch))
            rec = {'epoch': epoch, 'train_loss': self.train_epoch(train_loader, epoch, step)}
            last = epoch + 1 == num_epoch
            if work_dir and self.buckets.rank == 0 and ((epoch + 1) % save_interval == 0 or last):
                os.makedirs(work_dir, exist_ok=True)
                checkpoint.save_model(self.model, os.path.join(work_dir, 'epoch{}_model.pt'.format(epoch + 1)))
            if test_loader is not None and ((epoch + 1) % eval_interval == 0 or last):
                log('Eval epoch: {}'.format(epoch))
                rec['val_loss'], rec['acc'] = self.test_epoch(test_loader)
            history.append(rec)
        return history
