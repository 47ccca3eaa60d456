"""NumPy forms of the reference's augmentation helpers (feeder/tools.py:32-102), for the test phase
and for code that calls them directly.  The training hot path does NOT use them: it draws the same
random parameters (istgcn/pipeline.py, same generator calls in the same order) and applies the
transform on the GPU (istgcn_feeder_augment)."""
import numpy as np

from istgcn import pipeline


def _window(data_numpy, shift, size):
    C, T, V, M = data_numpy.shape
    if shift == 0 and size == T:
        return data_numpy
    if shift >= 0 and shift + size <= T:
        return data_numpy[:, shift:shift + size, :, :]
    out = np.zeros((C, size, V, M))
    out[:, -shift:-shift + T, :, :] = data_numpy
    return out


def auto_pading(data_numpy, size, random_pad=False):
    """tools.py:32-40: zero-pad a (C, T, V, M) clip shorter than ``size`` (at the front, or at a
    random offset)."""
    T = data_numpy.shape[1]
    if T >= size:
        return data_numpy
    shift, size = pipeline.draw_window(T, size, True) if random_pad else (0, size)
    return _window(data_numpy, shift, size)


def random_choose(data_numpy, size, auto_pad=True):
    """tools.py:43-56: a random ``size``-frame window (or random zero padding of a shorter clip)."""
    T = data_numpy.shape[1]
    if T < size and not auto_pad:
        return data_numpy
    shift, size = pipeline.draw_window(T, size, True)
    return _window(data_numpy, shift, size)


def random_move(data_numpy, angle_candidate=(-10., -5., 0., 5., 10.), scale_candidate=(0.9, 1.0, 1.1),
                transform_candidate=(-0.2, -0.1, 0.0, 0.1, 0.2), move_time_candidate=(1,)):
    """tools.py:59-102: per-frame rotation / scale / translation of the x, y channels, in place."""
    C, T, V, M = data_numpy.shape
    m = pipeline.draw_move(T, angle_candidate, scale_candidate, transform_candidate,
                           move_time_candidate).astype(np.float64)
    x, y = data_numpy[0].copy(), data_numpy[1].copy()
    ca, sa, tx, ty = (m[:, i].reshape(T, 1, 1) for i in range(4))
    data_numpy[0] = ca * x - sa * y + tx
    data_numpy[1] = sa * x + ca * y + ty
    return data_numpy
