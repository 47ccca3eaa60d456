"""Drop-in for the reference ``feeder.feeder.Feeder`` (feeder/feeder.py:21-85): memory-mapped
``(N, C, T, V, M)`` .npy + pickled ``(sample_name, label)``; same constructor arguments.

``device_augment=True`` (set by the processor for the training loader) makes ``__getitem__`` return
the RAW clip: the window / random_move parameters are then drawn per batch and applied on the GPU
by istgcn.pipeline.DevicePrefetcher (``augment_spec()`` hands it the feeder arguments)."""
import pickle

import numpy as np
import torch.utils.data

from . import tools
from istgcn.pipeline import AugmentSpec


class Feeder(torch.utils.data.Dataset):

    def __init__(self, data_path, label_path, random_choose=False, random_move=False, window_size=-1,
                 debug=False, mmap=True, device_augment=False):
        self.debug = debug
        self.data_path, self.label_path = data_path, label_path
        self.random_choose, self.random_move, self.window_size = random_choose, random_move, window_size
        self.device_augment = device_augment
        self.load_data(mmap)

    def load_data(self, mmap):
        with open(self.label_path, 'rb') as f:
            self.sample_name, self.label = pickle.load(f)
        self.data = np.load(self.data_path, mmap_mode='r') if mmap else np.load(self.data_path)
        if self.debug:
            self.label, self.data, self.sample_name = self.label[0:100], self.data[0:100], self.sample_name[0:100]
        self.N, self.C, self.T, self.V, self.M = self.data.shape

    def augment_spec(self):
        return AugmentSpec(self.random_choose, self.random_move, self.window_size)

    def __len__(self):
        return len(self.label)

    def __getitem__(self, index):
        data_numpy = np.array(self.data[index])
        label = self.label[index]
        if self.device_augment:
            return data_numpy, label
        if self.random_choose:
            data_numpy = tools.random_choose(data_numpy, self.window_size)
        elif self.window_size > 0:
            data_numpy = tools.auto_pading(data_numpy, self.window_size)
        if self.random_move:
            data_numpy = tools.random_move(data_numpy)
        return data_numpy, label
