from ctn_b200.models.tdcn import *  # noqa: F401,F403
from ctn_b200.models.tdcn import TimeDilatedConvNet, TimeDilatedConvBlock1d, ResidualBlock1d, DepthwiseSeparableConv1d  # noqa: F401
