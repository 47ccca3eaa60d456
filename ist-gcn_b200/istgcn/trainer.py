"""The training hot loop of the reference's REC_Processor (processor/recognition.py:185-310),
restated for one process per GPU: same initialisation (``weights_init`` :31-44), loss
(CrossEntropy :150), optimiser (SGD momentum 0.9, nesterov, weight decay :152-159) and step
schedule (``adjust_lr`` :168-176); the DataParallel wrap is replaced by ``dp.GradBuckets``.
Logging / TensorBoard / confusion matrices of the reference are out of scope.
"""
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
    for group in optimizer.param_groups:
        group['lr'] = lr
    return lr


class Trainer(object):
    """model + SGD(nesterov) + gradient buckets; ``step(x, label)`` is one iteration of
    recognition.py:249-298 (forward, loss, zero_grad, backward, optimizer step)."""

    def __init__(self, model, base_lr=0.1, weight_decay=1e-4, nesterov=True, momentum=0.9,
                 bucket_bytes=2 << 20, group=None):
        self.model = model
        self.buckets = dp.GradBuckets(list(model.named_parameters()), bucket_bytes, group)
        params = [p for b in self.buckets.buckets for _, p in b['params']]
        self.optimizer = torch.optim.SGD(params, lr=base_lr, momentum=momentum, nesterov=nesterov,
                                         weight_decay=weight_decay, foreach=True)
        self.base_lr = base_lr

    def step(self, x, label):
        self.model.train()
        self.buckets.zero()
        output = self.model(x)
        loss = F.cross_entropy(output, label)
        loss.backward()
        self.buckets.finish()
        self.optimizer.step()
        return loss

    @torch.no_grad()
    def evaluate(self, x):
        self.model.eval()
        return self.model(x)
