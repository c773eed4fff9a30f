"""choose_nonlinear mirroring src/utils/model.py:3-20 (parameter-free activations are plain torch modules)."""
import torch.nn as nn


def choose_nonlinear(name, **kwargs):
    if name == 'relu':
        return nn.ReLU()
    if name == 'sigmoid':
        return nn.Sigmoid()
    if name == 'softmax':
        assert 'dim' in kwargs, "dim is expected for softmax."
        return nn.Softmax(**kwargs)
    if name == 'tanh':
        return nn.Tanh()
    if name == 'leaky-relu':
        return nn.LeakyReLU()
    if name == 'gelu':
        return nn.GELU()
    raise NotImplementedError("Invalid nonlinear function is specified. Choose 'relu' instead of {}.".format(name))
