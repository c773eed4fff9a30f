from ctn_b200.criterion.sdr import sisdr, SISDR, NegSISDR, EPS  # noqa: F401
