from . import dnn_to_bnn  # noqa: F401
