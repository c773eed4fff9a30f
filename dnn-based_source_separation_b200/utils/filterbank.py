from ctn_b200.utils.filterbank import choose_filterbank  # noqa: F401
