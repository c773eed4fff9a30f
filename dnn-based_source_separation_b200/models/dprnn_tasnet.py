from ctn_b200.models.dprnn_tasnet import *  # noqa: F401,F403
