from ctn_b200.models.dprnn import *  # noqa: F401,F403
