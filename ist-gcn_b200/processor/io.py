"""Argument / environment / checkpoint plumbing of the reference's processors.

reference: processor/my_io.py:31-87 (``load_arg``: argparse defaults <- YAML <- command line with
an assertion on unknown YAML keys; ``init_environment``; ``load_weights``; ``gpu``),
torchlight/torchlight/io.py:101-130 (``save_model``, ``save_arg`` -> ``config.yaml``,
``print_log`` -> stdout + ``log.txt``) and :160-203 (``str2bool``, ``DictAction``).

What changes: ``gpu()``.  The reference wraps the model in a single-process ``nn.DataParallel`` over
``--device``; here every GPU is one process (launch with ``torchrun --nproc-per-node N main.py
recognition -c ...``; RANK / LOCAL_RANK / WORLD_SIZE from the environment), gradients are averaged by
istgcn.dp.GradBuckets over NCCL and BatchNorm statistics stay per rank, which is what DataParallel's
replicas compute.  Without torchrun, ``--device`` selects the one GPU to use."""
import argparse
import os
import sys
import time

import torch
import torch.distributed as dist
import yaml

from istgcn import checkpoint


def str2bool(v):
    if v.lower() in ('yes', 'true', 't', 'y', '1'):
        return True
    if v.lower() in ('no', 'false', 'f', 'n', '0'):
        return False
    raise argparse.ArgumentTypeError('Boolean value expected.')


class DictAction(argparse.Action):
    """``--model_args "dropout=0.5, num_class=60"``: merged into the YAML / default dict
    (torchlight/io.py:192-203; the reference evaluates the string as ``dict(...)`` too)."""

    def __init__(self, option_strings, dest, nargs=None, **kwargs):
        if nargs is not None:
            raise ValueError('nargs not allowed')
        super().__init__(option_strings, dest, **kwargs)

    def __call__(self, parser, namespace, values, option_string=None):
        merged = dict(getattr(namespace, self.dest) or {})
        merged.update(eval('dict({})'.format(values)))     # noqa: S307 -- reference semantics
        setattr(namespace, self.dest, merged)


class IO(object):
    """Base of the processors: arguments, work directory, model, weights, device."""

    def __init__(self, argv=None):
        self.load_arg(argv)
        self.init_environment()
        self.load_model()
        self.load_weights()
        self.gpu()

    # ---- my_io.py:31-50
    def load_arg(self, argv=None):
        parser = self.get_parser()
        p = parser.parse_args(argv)
        if p.config is not None:
            with open(p.config, 'r', encoding='utf-8') as f:
                default_arg = yaml.load(f, Loader=yaml.FullLoader) or {}
            known = vars(p).keys()
            for k in default_arg.keys():
                if k not in known:
                    print('Unknown Arguments: {}'.format(k))
                    assert k in known
            parser.set_defaults(**default_arg)
        self.arg = parser.parse_args(argv)

    # ---- my_io.py:52-66 + torchlight/io.py:109-130
    def init_environment(self):
        self.world = int(os.environ.get('WORLD_SIZE', '1'))
        self.rank = int(os.environ.get('RANK', '0'))
        self.local_rank = int(os.environ.get('LOCAL_RANK', '0'))
        self.work_dir = self.arg.work_dir
        if self.rank == 0:
            os.makedirs(self.work_dir, exist_ok=True)
            with open(os.path.join(self.work_dir, 'config.yaml'), 'w') as f:
                f.write('# command line: {}\n\n'.format(' '.join(sys.argv)))
                yaml.dump(vars(self.arg), f, default_flow_style=False, indent=4)
        if self.arg.use_gpu:
            devices = self.arg.device if isinstance(self.arg.device, (list, tuple)) else [self.arg.device]
            if self.world > 1:
                index = self.local_rank
            else:
                index = int(devices[0])
                if len(devices) > 1:
                    self.print_log('{} devices requested: start one process per GPU with `torchrun '
                                   '--nproc-per-node {}` (single-process DataParallel is not used); '
                                   'running on cuda:{}'.format(len(devices), len(devices), index))
            if not torch.cuda.is_available():
                raise RuntimeError('istgcn_b200: use_gpu=True but no CUDA device is visible (there is no CPU path)')
            torch.cuda.set_device(index)
            self.dev = torch.device('cuda', index)
            if self.world > 1 and not dist.is_initialized():
                dist.init_process_group('nccl', device_id=self.dev)
        else:
            raise RuntimeError('istgcn_b200: the models only run on a CUDA sm_100a device (use_gpu=False is not supported)')

    def print_log(self, msg, print_time=True):
        if print_time:
            msg = time.strftime('[%m.%d.%y|%X] ', time.localtime()) + msg
        if self.rank != 0:
            return
        if self.arg.print_log:
            print(msg)
        if self.arg.save_log:
            with open(os.path.join(self.work_dir, 'log.txt'), 'a') as f:
                print(msg, file=f)

    # ---- my_io.py:68-75, torchlight/io.py:51-107
    def load_model(self):
        self.model = checkpoint.load_model(self.arg.model, **(self.arg.model_args))

    def load_weights(self):
        if self.arg.weights:
            self.model = checkpoint.load_weights(self.model, self.arg.weights, self.arg.ignore_weights,
                                                 log=self.print_log)

    def save_model(self, model, name):
        if self.rank == 0:
            path = '{}/{}'.format(self.work_dir, name)
            checkpoint.save_model(model, path)
            self.print_log('The model has been saved as {}.'.format(path))

    # ---- my_io.py:77-87
    def gpu(self):
        self.model = self.model.to(self.dev)
        for name, value in vars(self).items():
            if isinstance(value, torch.nn.Module) and name != 'model':
                setattr(self, name, value.to(self.dev))
        if self.world > 1:
            from istgcn import dp
            dp.broadcast_state(self.model)

    def start(self):
        self.print_log('Parameters:\n{}\n'.format(str(vars(self.arg))))

    @staticmethod
    def get_parser(add_help=False):
        parser = argparse.ArgumentParser(add_help=add_help, description='IO Processor')
        parser.add_argument('-w', '--work_dir', default='./work_dir/tmp', help='the work folder for storing results')
        parser.add_argument('-c', '--config', default=None, help='path to the configuration file')
        parser.add_argument('--use_gpu', type=str2bool, default=True, help='use GPUs or not')
        parser.add_argument('--device', type=int, default=0, nargs='+', help='the indexes of GPUs for training or testing')
        parser.add_argument('--print_log', type=str2bool, default=True, help='print logging or not')
        parser.add_argument('--save_log', type=str2bool, default=True, help='save logging or not')
        parser.add_argument('--model', default=None, help='the model will be used')
        parser.add_argument('--model_args', action=DictAction, default=dict(), help='the arguments of model')
        parser.add_argument('--weights', default=None, help='the weights for network initialization')
        parser.add_argument('--ignore_weights', type=str, default=[], nargs='+',
                            help='the name of weights which will be ignored in the initialization')
        return parser
