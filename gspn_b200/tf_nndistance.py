"""Drop-in for the reference module of the same name (tf_ops/*/tf_nndistance.py)."""
from .ops import nn_distance  # noqa: F401
