from ctn_b200.models.conv_tasnet import *  # noqa: F401,F403
from ctn_b200.models.conv_tasnet import ConvTasNet, Separator  # noqa: F401
