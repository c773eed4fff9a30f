from ctn_b200.modules.conv import *  # noqa: F401,F403
