from ctn_b200.models.tcn import TemporalConvNet, ConvBlock1d, ResidualBlock1d, DepthwiseSeparableConv1d  # noqa: F401
