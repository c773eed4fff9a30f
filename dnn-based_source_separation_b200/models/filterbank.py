from ctn_b200.models.filterbank import Encoder, Decoder  # noqa: F401
