"""Deprecated alias module, mirroring src/models/tcn.py of the reference (``TemporalConvNet`` == ``TimeDilatedConvNet``,
``ConvBlock1d`` == ``TimeDilatedConvBlock1d``; tcn.py:9,19-23).  Same kernels, same state_dict keys."""
import warnings

from .tdcn import (TimeDilatedConvNet, TimeDilatedConvBlock1d, ResidualBlock1d, DepthwiseSeparableConv1d, EPS)  # noqa: F401

warnings.warn("Use models.tdcn instead.", FutureWarning)


class TemporalConvNet(TimeDilatedConvNet):
    def __init__(self, *args, **kwargs):
        warnings.warn("Use TimeDilatedConvNet instead.", DeprecationWarning)
        super().__init__(*args, **kwargs)


class ConvBlock1d(TimeDilatedConvBlock1d):
    pass
