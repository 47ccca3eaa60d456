"""Drop-in for the reference ``processor.recognition.REC_Processor`` (+ its base ``Processor``):
``python main.py recognition -c config/.../train.yaml [--overrides]``.

reference: processor/processor.py:32-226 (data loaders, epoch loop, save / eval cadence),
processor/recognition.py:146-183 (model init, optimiser, step LR, top-k), :185-310 (train),
:312-385 (test), the two argument parsers (:387-410, processor.py:228-278).  Same YAML keys, same
command line, same ``epoch{N}_model.pt`` / ``config.yaml`` / ``log.txt`` artefacts.

Kept out (not part of the hot path): TensorBoard scalars, matplotlib plots, the
confusion-matrix / loss-acc CSV files and the end-of-run ``os.rename`` of the work directory.

The iteration itself is istgcn.trainer.Trainer.step (CUDA graph, flat-bucket SGD kernel, NCCL
gradient buckets); batches reach the GPU through istgcn.pipeline.DevicePrefetcher (pinned double
buffer + copy stream, augmentation on the GPU).  The reference reads the loss back every iteration
(``loss.data.item()``, :292, a host sync); here it is read every ``log_interval`` iterations."""
import argparse
import time

import numpy as np
import torch
import torch.nn.functional as F

from istgcn import checkpoint, pipeline, trainer

from .io import IO, DictAction, str2bool


class Processor(IO):

    def __init__(self, argv=None):
        self.load_arg(argv)
        self.init_environment()
        self.load_model()
        self.load_weights()
        self.gpu()
        self.load_data()
        self.load_optimizer()

    def init_environment(self):
        super().init_environment()
        self.result = dict()
        self.iter_info = dict()
        self.epoch_info = dict()
        self.meta_info = dict(epoch=0, iter=0)

    # ---- processor.py:60-83
    def load_data(self):
        Feeder = checkpoint.import_class(self.arg.feeder)
        if 'debug' not in self.arg.train_feeder_args:
            self.arg.train_feeder_args['debug'] = self.arg.debug
        self.data_loader = dict()
        self.augment = None
        if self.arg.phase == 'train':
            args = dict(self.arg.train_feeder_args)
            try:                                    # our feeder: raw clips, augmentation on the GPU
                dataset = Feeder(device_augment=True, **args)
                self.augment = dataset.augment_spec()
            except TypeError:                       # a user-supplied feeder without that switch
                dataset = Feeder(**args)
            sampler = None
            if self.world > 1:
                sampler = torch.utils.data.distributed.DistributedSampler(
                    dataset, num_replicas=self.world, rank=self.rank, shuffle=True, drop_last=True)
            self.train_sampler = sampler
            self.data_loader['train'] = torch.utils.data.DataLoader(
                dataset=dataset, batch_size=self.arg.batch_size, shuffle=sampler is None, sampler=sampler,
                num_workers=self.arg.num_worker, drop_last=True)
        if self.arg.test_feeder_args:
            self.data_loader['test'] = torch.utils.data.DataLoader(
                dataset=Feeder(**self.arg.test_feeder_args), batch_size=self.arg.test_batch_size,
                shuffle=False, num_workers=self.arg.num_worker)

    def load_optimizer(self):
        pass

    # ---- processor.py:85-118
    def show_epoch_info(self):
        for k, v in self.epoch_info.items():
            self.print_log('\t{}: {}'.format(k, v))

    def show_iter_info(self):
        if self.meta_info['iter'] % self.arg.log_interval == 0:
            info = '\tIter {} Done.'.format(self.meta_info['iter'])
            for k, v in self.iter_info.items():
                info += ' | {}: {:.4f}'.format(k, v) if isinstance(v, float) else ' | {}: {}'.format(k, v)
            self.print_log(info)

    # ---- processor.py:159-226
    def start(self):
        self.print_log('Parameters:\n{}\n'.format(str(vars(self.arg))))
        if self.arg.phase == 'train':
            for epoch in range(self.arg.start_epoch, self.arg.num_epoch):
                self.meta_info['epoch'] = epoch
                self.print_log('Training epoch: {}'.format(epoch))
                self.train()
                self.print_log('Done.')
                last = epoch + 1 == self.arg.num_epoch
                if (epoch + 1) % self.arg.save_interval == 0 or last:
                    self.save_model(self.model, 'epoch{}_model.pt'.format(epoch + 1))
                if ((epoch + 1) % self.arg.eval_interval == 0 or last) and 'test' in self.data_loader:
                    self.print_log('Eval epoch: {}'.format(epoch))
                    self.test()
                    self.print_log('Done.')
        elif self.arg.phase == 'test':
            if self.arg.weights is None:
                raise ValueError('Please appoint --weights.')
            self.print_log('Model:   {}.'.format(self.arg.model))
            self.print_log('Weights: {}.'.format(self.arg.weights))
            self.print_log('Evaluation Start:')
            self.test()
            self.print_log('Done.\n')
        if self.world > 1:
            torch.distributed.barrier()

    @staticmethod
    def get_parser(add_help=False):
        parser = argparse.ArgumentParser(add_help=add_help, description='Base Processor')
        parser.add_argument('-w', '--work_dir', default='./work_dir/tmp', help='the work folder for storing results')
        parser.add_argument('-c', '--config', default=None, help='path to the configuration file')
        parser.add_argument('--phase', default='train', help='must be train or test')
        parser.add_argument('--save_result', type=str2bool, default=False,
                            help='if ture, the output of the model will be stored')
        parser.add_argument('--start_epoch', type=int, default=0, help='start training from which epoch')
        parser.add_argument('--num_epoch', type=int, default=80, help='stop training in which epoch')
        parser.add_argument('--use_gpu', type=str2bool, default=True, help='use GPUs or not')
        parser.add_argument('--device', type=int, default=0, nargs='+',
                            help='the indexes of GPUs for training or testing')
        parser.add_argument('--log_interval', type=int, default=100,
                            help='the interval for printing messages (#iteration)')
        parser.add_argument('--save_interval', type=int, default=10,
                            help='the interval for storing models (#iteration)')
        parser.add_argument('--eval_interval', type=int, default=5,
                            help='the interval for evaluating models (#iteration)')
        parser.add_argument('--save_log', type=str2bool, default=True, help='save logging or not')
        parser.add_argument('--print_log', type=str2bool, default=True, help='print logging or not')
        parser.add_argument('--pavi_log', type=str2bool, default=False, help='logging on pavi or not')
        parser.add_argument('--feeder', default='feeder.feeder', help='data loader will be used')
        parser.add_argument('--num_worker', type=int, default=0, help='the number of worker per gpu for data loader')
        parser.add_argument('--train_feeder_args', action=DictAction, default=dict(),
                            help='the arguments of data loader for training')
        parser.add_argument('--test_feeder_args', action=DictAction, default=dict(),
                            help='the arguments of data loader for test')
        parser.add_argument('--batch_size', type=int, default=256, help='training batch size')
        parser.add_argument('--test_batch_size', type=int, default=256, help='test batch size')
        parser.add_argument('--debug', action='store_true', help='less data, faster loading')
        parser.add_argument('--model', default=None, help='the model will be used')
        parser.add_argument('--model_args', action=DictAction, default=dict(), help='the arguments of model')
        parser.add_argument('--weights', default=None, help='the weights for network initialization')
        parser.add_argument('--ignore_weights', type=str, default=[], nargs='+',
                            help='the name of weights which will be ignored in the initialization')
        return parser


class REC_Processor(Processor):
    """Processor for skeleton-based action recognition."""

    # ---- recognition.py:146-150
    def load_model(self):
        self.model = checkpoint.load_model(self.arg.model, **(self.arg.model_args))
        self.model.apply(trainer.weights_init)

    # ---- recognition.py:152-166
    def load_optimizer(self):
        if self.arg.optimizer != 'SGD':
            raise ValueError('istgcn_b200: only optimizer=SGD (the reference configs\' choice, '
                             'recognition.py:153-159) is implemented on the flat-bucket kernel')
        self.trainer = trainer.Trainer(self.model, base_lr=self.arg.base_lr, weight_decay=self.arg.weight_decay,
                                       nesterov=self.arg.nesterov, momentum=0.9,
                                       use_graph=not getattr(self.arg, 'no_graph', False))
        self.optimizer = self.trainer.optimizer

    # ---- recognition.py:168-176
    def adjust_lr(self):
        if self.arg.optimizer == 'SGD' and self.arg.step:
            lr = self.arg.base_lr * (0.1 ** np.sum(self.meta_info['epoch'] >= np.array(self.arg.step)))
        else:
            lr = self.arg.base_lr
        self.trainer.set_lr(float(lr))
        self.lr = float(lr)

    # ---- recognition.py:178-183
    def show_topk(self, k):
        rank = self.result.argsort()
        hit_top_k = [l in rank[i, -k:] for i, l in enumerate(self.label)]
        accuracy = sum(hit_top_k) * 1.0 / len(hit_top_k)
        self.print_log('\tTop{}: {:.2f}%'.format(k, 100 * accuracy))
        return accuracy

    # ---- recognition.py:185-310 (the hot loop is :249-298)
    def train(self):
        self.adjust_lr()
        sampler = getattr(self, 'train_sampler', None)
        if sampler is not None:
            sampler.set_epoch(self.meta_info['epoch'])
        loader = pipeline.DevicePrefetcher(self.data_loader['train'], self.dev, self.augment)
        total, count, t0 = None, 0, time.time()
        for data, label in loader:
            loss = self.trainer.step(data, label).detach()
            total = loss.clone() if total is None else total + loss
            count += 1
            self.meta_info['iter'] += 1
            if self.meta_info['iter'] % self.arg.log_interval == 0:
                self.iter_info['loss'] = loss.item()           # the only per-iteration host sync
                self.iter_info['lr'] = '{:.6f}'.format(self.lr)
                self.show_iter_info()
        self.epoch_info['mean_loss'] = float(total.item()) / count if count else float('nan')
        self.epoch_info['clips_per_s'] = '{:.1f}'.format(
            count * self.arg.batch_size * self.world / max(time.time() - t0, 1e-9))
        self.show_epoch_info()

    # ---- recognition.py:312-385
    @torch.no_grad()
    def test(self, evaluation=True):
        self.model.eval()
        result_frag, label_frag, loss_value = [], [], []
        for data, label in self.data_loader['test']:
            data = data.float().to(self.dev, non_blocking=True)
            label = label.long().to(self.dev, non_blocking=True)
            output = self.model(data)
            result_frag.append(output.cpu().numpy())
            if evaluation:
                loss_value.append(F.cross_entropy(output, label).item())
                label_frag.append(label.cpu().numpy())
        self.result = np.concatenate(result_frag)
        if evaluation:
            self.label = np.concatenate(label_frag)
            self.epoch_info['mean_loss'] = float(np.mean(loss_value))
            self.show_epoch_info()
            for k in self.arg.show_topk:
                self.show_topk(k)

    @staticmethod
    def get_parser(add_help=False):
        parent_parser = Processor.get_parser(add_help=False)
        parser = argparse.ArgumentParser(add_help=add_help, parents=[parent_parser],
                                         description='Spatial Temporal Graph Convolution Network')
        parser.add_argument('--show_topk', type=int, default=[1, 5], nargs='+',
                            help='which Top K accuracy will be shown')
        parser.add_argument('--base_lr', type=float, default=0.01, help='initial learning rate')
        parser.add_argument('--step', type=int, default=[], nargs='+',
                            help='the epoch where optimizer reduce the learning rate')
        parser.add_argument('--optimizer', default='SGD', help='type of optimizer')
        parser.add_argument('--nesterov', type=str2bool, default=True, help='use nesterov or not')
        parser.add_argument('--weight_decay', type=float, default=0.0001, help='weight decay for optimizer')
        parser.add_argument('--no_graph', type=str2bool, default=False,
                            help='launch the kernels eagerly instead of replaying a CUDA graph')
        return parser
