"""Drop-in for the reference module of the same name (tf_ops/*/tf_interpolate.py)."""
from .ops import three_nn, three_interpolate  # noqa: F401
