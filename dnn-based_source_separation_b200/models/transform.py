from ctn_b200.models.transform import *  # noqa: F401,F403
