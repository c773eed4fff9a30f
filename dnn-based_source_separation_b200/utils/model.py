from ctn_b200.utils.model import choose_nonlinear  # noqa: F401
