"""NumPy <-> CUDA-tensor interchange for the drop-in layer: results come back in the caller's kind."""
import numpy as np
import torch

from . import engine


def is_numpy(a):
    return isinstance(a, np.ndarray)


def as_cuda(a, dtype=None, device=None):
    return engine.to_device(a, device=device, dtype=dtype)


def give_back(t, want_numpy):
    if want_numpy and isinstance(t, torch.Tensor):
        return t.cpu().numpy()
    return t
