from ctn_b200.modules.norm import GlobalLayerNorm, CumulativeLayerNorm1d, EPS  # noqa: F401
