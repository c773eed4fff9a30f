from ctn_b200.utils.tasnet import choose_layer_norm  # noqa: F401
