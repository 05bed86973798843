"""`BaseVariationalLayer_` and `get_kernel_size` (reference: layers/base_variational_layer.py:35-68)."""
from .._base import BaseVariationalLayer_, get_kernel_size

__all__ = ["BaseVariationalLayer_", "get_kernel_size"]
